// gdb_pack.cpp -- host-side octile packer (pure C++, no CUDA).
//
// Turns one graph (node AoS + undirected edge list) into the
// position-independent device blob described in include/graphdot_b200.h.
// Replaces the numpy packer of reference
// graphdot/kernel/marginalized/_octilegraph.py:100-177 (degrees :113-139,
// symmetric replication :142-147, tile sort :150-158, bit masks :160-168).
// Differences by design: octiles are sorted by (tile row, tile column) and
// carry one row-major mask, elements are stored row-major inside a tile and a
// per-tile-row index (CSR over tiles) is added, because the B200 matvec
// gathers along rows instead of scattering with atomics.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gdb_internal.h"

namespace {

inline uint64_t pad16(uint64_t v) { return (v + 15u) & ~uint64_t(15); }

struct Sizes {
    uint32_t edge_size, label_off;
    uint32_t nnz, n_octile, n_tile;
    uint64_t off_degree, off_node, off_octile, off_tilerow, off_edge, off_pool, total;
    uint64_t off_emeta, off_rowptr, off_rowadj, off_tileelem, off_ellslot, off_lanemap;
    uint64_t off_tcptr, off_tccol, off_tcslot;
    uint32_t n_tc;
};

struct Nz {
    uint32_t key;  // (trow, tcol, row, col) packed: trow<<19 | tcol<<6 | row<<3 | col
    uint32_t i, j, e;
    uint32_t ord;  // position in the reference's [edges ; swapped edges] list (duplicate resolution)
};

bool collect(const gdb_graph_src *g, std::vector<Nz> &nz) {
    nz.clear();
    nz.reserve(2 * (size_t)g->n_edge);
    for (uint32_t k = 0; k < g->n_edge; ++k) {
        const uint32_t i = g->edge_i[k], j = g->edge_j[k];
        if (i >= g->n_node || j >= g->n_node) return false;
        nz.push_back({0, i, j, k, k});
        if (i != j) nz.push_back({0, j, i, k, g->n_edge + k});
    }
    for (auto &z : nz) z.key = ((z.i >> 3) << 19) | ((z.j >> 3) << 6) | ((z.i & 7) << 3) | (z.j & 7);
    // duplicate directed entries (multi-edges): keep the FIRST of the list
    // [edges ; swapped edges], like the reference's np.unique(..., return_index=True)
    // (_octilegraph.py:141-158)
    std::sort(nz.begin(), nz.end(), [](const Nz &a, const Nz &b) { return a.key != b.key ? a.key < b.key : a.ord < b.ord; });
    size_t w = 0;
    for (size_t r = 0; r < nz.size(); ++r)
        if (w == 0 || nz[w - 1].key != nz[r].key) nz[w++] = nz[r];
    nz.resize(w);
    return true;
}

void plan(const gdb_layout *L, const gdb_graph_src *g, const std::vector<Nz> &nz, Sizes &s) {
    if (L->weighted) {
        const uint32_t a = std::max<uint32_t>(4u, L->edge_label_align);
        s.label_off = (4u + L->edge_label_align - 1u) / L->edge_label_align * L->edge_label_align;
        s.edge_size = (s.label_off + L->edge_label_size + a - 1u) / a * a;
    } else {
        s.label_off = 0;
        s.edge_size = L->edge_label_size;
    }
    s.nnz = (uint32_t)nz.size();
    s.n_tile = (g->n_node + 7u) / 8u;
    uint32_t n_oct = 0;
    for (size_t k = 0; k < nz.size(); ++k)
        if (k == 0 || (nz[k].key >> 6) != (nz[k - 1].key >> 6)) ++n_oct;
    s.n_octile = n_oct;
    s.off_degree = GDB_HDR_BYTES;
    s.off_node = s.off_degree + pad16(4ull * g->n_node);
    s.off_octile = s.off_node + pad16((uint64_t)L->node_size * g->n_node);
    s.off_tilerow = s.off_octile + 16ull * s.n_octile;
    s.off_edge = s.off_tilerow + pad16(4ull * (s.n_tile + 1));
    // row index derived from the octiles (pair-independent, so built once
    // here instead of per pair on the device): per element (row | col << 16)
    // in octile order, CSR over rows with (col | element << 16), and the
    // first element of every tile row
    s.off_emeta = s.off_edge + pad16((uint64_t)s.edge_size * s.nnz);
    s.off_rowptr = s.off_emeta + pad16(4ull * s.nnz);
    s.off_rowadj = s.off_rowptr + pad16(4ull * (g->n_node + 1));
    s.off_tileelem = s.off_rowadj + pad16(4ull * s.nnz);
    s.off_ellslot = s.off_tileelem + pad16(4ull * (s.n_tile + 1));
    s.off_lanemap = s.off_ellslot + pad16(4ull * s.nnz);
    // distinct columns per tile row: nz is sorted by (tile row, tile col, row, col)
    {
        std::vector<uint32_t> cols;
        uint32_t n_tc = 0;
        size_t k = 0;
        while (k < nz.size()) {
            const uint32_t t = nz[k].i >> 3;
            cols.clear();
            for (; k < nz.size() && (nz[k].i >> 3) == t; ++k) cols.push_back(nz[k].j);
            std::sort(cols.begin(), cols.end());
            n_tc += (uint32_t)(std::unique(cols.begin(), cols.end()) - cols.begin());
        }
        s.n_tc = n_tc;
    }
    s.off_tcptr = s.off_lanemap + pad16(4ull * g->n_node);
    s.off_tccol = s.off_tcptr + pad16(4ull * (s.n_tile + 1));
    s.off_tcslot = s.off_tccol + pad16(2ull * s.n_tc);
    s.off_pool = s.off_tcslot + pad16(2ull * s.nnz);
    s.total = s.off_pool + pad16(g->pool_bytes);
}

}  // namespace

extern "C" int gdb_graph_packed_size(const gdb_layout *layout, const gdb_graph_src *g, uint64_t *bytes) {
    if (!layout || !g || !bytes) return gdb_fail(GDB_ERR_INVALID, "gdb_graph_packed_size: null argument");
    if (g->n_node == 0 || g->n_node >= (1u << 16)) return gdb_fail(GDB_ERR_INVALID, "graph must have 1..65535 nodes");
    std::vector<Nz> nz;
    if (!collect(g, nz)) return gdb_fail(GDB_ERR_INVALID, "edge end point out of range");
    Sizes s;
    plan(layout, g, nz, s);
    if (s.total >= (1ull << 32)) return gdb_fail(GDB_ERR_INVALID, "packed graph exceeds 4 GiB");
    *bytes = s.total;
    return GDB_OK;
}

extern "C" int gdb_graph_pack(const gdb_layout *L, const gdb_graph_src *g, void *blob, uint64_t capacity) {
    if (!L || !g || !blob) return gdb_fail(GDB_ERR_INVALID, "gdb_graph_pack: null argument");
    std::vector<Nz> nz;
    if (!collect(g, nz)) return gdb_fail(GDB_ERR_INVALID, "edge end point out of range");
    Sizes s;
    plan(L, g, nz, s);
    if (s.total > capacity) return gdb_fail(GDB_ERR_INVALID, "gdb_graph_pack: blob too small");
    uint8_t *base = static_cast<uint8_t *>(blob);
    std::memset(base, 0, s.total);

    gdb_graph_hdr_host *h = reinterpret_cast<gdb_graph_hdr_host *>(base);
    h->n_node = (int32_t)g->n_node;
    h->n_octile = (int32_t)s.n_octile;
    h->nnz = (int32_t)s.nnz;
    h->n_tile = (int32_t)s.n_tile;
    h->off_degree = (uint32_t)s.off_degree;
    h->off_node = (uint32_t)s.off_node;
    h->off_octile = (uint32_t)s.off_octile;
    h->off_tilerow = (uint32_t)s.off_tilerow;
    h->off_edge = (uint32_t)s.off_edge;
    h->off_pool = (uint32_t)s.off_pool;
    h->blob_bytes = (uint32_t)s.total;
    h->flags = (L->weighted ? 1u : 0u) | (s.nnz < 65536u ? 2u : 0u);  // bit 1: 16-bit row index valid
    h->off_emeta = (uint32_t)s.off_emeta;
    h->off_rowptr = (uint32_t)s.off_rowptr;
    h->off_rowadj = (uint32_t)s.off_rowadj;
    h->off_tileelem = (uint32_t)s.off_tileelem;

    // degrees: sum of incident weights, self loops once, 0 -> 1
    std::vector<double> deg(g->n_node, 0.0);
    for (uint32_t k = 0; k < g->n_edge; ++k) {
        const double w = g->edge_w ? (double)g->edge_w[k] : 1.0;
        deg[g->edge_i[k]] += w;
        if (g->edge_i[k] != g->edge_j[k]) deg[g->edge_j[k]] += w;
    }
    float *degree = reinterpret_cast<float *>(base + s.off_degree);
    for (uint32_t i = 0; i < g->n_node; ++i) degree[i] = deg[i] == 0.0 ? 1.0f : (float)deg[i];

    // nodes: verbatim AoS; frozen_array slots become blob-relative offsets
    std::memcpy(base + s.off_node, g->nodes, (size_t)L->node_size * g->n_node);
    for (uint32_t i = 0; i < g->n_node; ++i)
        for (uint32_t f = 0; f < L->n_node_ptr; ++f) {
            uint64_t *slot = reinterpret_cast<uint64_t *>(base + s.off_node + (size_t)i * L->node_size + L->node_ptr_offset[f]);
            *slot += s.off_pool;
        }

    // octiles, tile rows, elements
    gdb_octile_host *oct = reinterpret_cast<gdb_octile_host *>(base + s.off_octile);
    uint32_t *tilerow = reinterpret_cast<uint32_t *>(base + s.off_tilerow);
    uint8_t *edges = base + s.off_edge;
    int32_t o = -1;
    for (uint32_t k = 0; k < s.nnz; ++k) {
        const Nz &z = nz[k];
        if (k == 0 || (z.key >> 6) != (nz[k - 1].key >> 6)) {
            ++o;
            oct[o].mask = 0;
            oct[o].start = k;
            oct[o].trow = (uint16_t)(z.i >> 3);
            oct[o].tcol = (uint16_t)(z.j >> 3);
        }
        oct[o].mask |= 1ull << (z.key & 63u);
        uint8_t *e = edges + (size_t)k * s.edge_size;
        if (L->weighted) {
            const float w = g->edge_w ? g->edge_w[z.e] : 1.0f;
            std::memcpy(e, &w, 4);
        }
        if (L->edge_label_size)
            std::memcpy(e + s.label_off, static_cast<const uint8_t *>(g->edge_labels) + (size_t)z.e * L->edge_label_size,
                        L->edge_label_size);
        for (uint32_t f = 0; f < L->n_edge_ptr; ++f) {
            uint64_t *slot = reinterpret_cast<uint64_t *>(e + s.label_off + L->edge_ptr_offset[f]);
            *slot += s.off_pool;
        }
    }
    // row index
    {
        uint32_t *emeta = reinterpret_cast<uint32_t *>(base + s.off_emeta);
        uint32_t *rowptr = reinterpret_cast<uint32_t *>(base + s.off_rowptr);
        uint32_t *rowadj = reinterpret_cast<uint32_t *>(base + s.off_rowadj);
        uint32_t *tileelem = reinterpret_cast<uint32_t *>(base + s.off_tileelem);
        std::vector<uint32_t> fill(g->n_node + 1, 0);
        for (uint32_t k = 0; k < s.nnz; ++k) {
            emeta[k] = (nz[k].i & 0xffffu) | (nz[k].j << 16);
            fill[nz[k].i + 1]++;
        }
        uint32_t max_degree = 0;
        for (uint32_t i = 0; i < g->n_node; ++i) max_degree = std::max(max_degree, fill[i + 1]);
        h->max_degree = max_degree;
        h->off_ellslot = (uint32_t)s.off_ellslot;
        for (uint32_t i = 0; i < g->n_node; ++i) fill[i + 1] += fill[i];
        for (uint32_t i = 0; i <= g->n_node; ++i) rowptr[i] = fill[i];
        {
            // lane map of the small-pair kernel: nodes in order of decreasing degree
            // (stable).  lanemap[p] & 0xffff = node at position p, lanemap[i] >> 16 =
            // position of node i.  rowpos[k] = row | (index within the row << 16) of
            // CSR element k.  vcols = number of virtual columns (chunks of 2 / of 4
            // neighbour slots) the kernel needs lanes for.
            uint32_t *rowpos = reinterpret_cast<uint32_t *>(base + s.off_ellslot);
            uint32_t *lanemap = reinterpret_cast<uint32_t *>(base + s.off_lanemap);
            h->off_lanemap = (uint32_t)s.off_lanemap;
            std::vector<uint32_t> order(g->n_node);
            for (uint32_t i = 0; i < g->n_node; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                return rowptr[a + 1] - rowptr[a] > rowptr[b + 1] - rowptr[b];
            });
            for (uint32_t pos = 0; pos < g->n_node; ++pos) lanemap[pos] = order[pos];
            for (uint32_t pos = 0; pos < g->n_node; ++pos) lanemap[order[pos]] |= pos << 16;
            uint32_t nv2 = 0, nv4 = 0;
            for (uint32_t i = 0; i < g->n_node; ++i) {
                const uint32_t deg = rowptr[i + 1] - rowptr[i];
                nv2 += std::max(1u, (deg + 1u) / 2u);
                nv4 += std::max(1u, (deg + 3u) / 4u);
                for (uint32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) rowpos[k] = i | ((k - rowptr[i]) << 16);
            }
            h->vcols = std::min(nv2, 0xffffu) | (std::min(nv4, 0xffffu) << 16);
        }
        // nz is sorted by (tile row, tile col, row, col): filling in this order
        // leaves every row's neighbours sorted by column
        for (uint32_t k = 0; k < s.nnz; ++k) rowadj[fill[nz[k].i]++] = (nz[k].j & 0xffffu) | (k << 16);
        // neighbour-row lists per tile row and the slot of every CSR element in them
        {
            uint32_t *tcptr = reinterpret_cast<uint32_t *>(base + s.off_tcptr);
            uint16_t *tccol = reinterpret_cast<uint16_t *>(base + s.off_tccol);
            uint16_t *tcslot = reinterpret_cast<uint16_t *>(base + s.off_tcslot);
            h->off_tcptr = (uint32_t)s.off_tcptr;
            h->off_tccol = (uint32_t)s.off_tccol;
            h->off_tcslot = (uint32_t)s.off_tcslot;
            std::vector<uint32_t> cols;
            uint32_t at = 0, max_tc = 0;
            for (uint32_t t = 0; t < s.n_tile; ++t) {
                tcptr[t] = at;
                const uint32_t r0 = t * 8u, r1 = std::min(g->n_node, r0 + 8u);
                cols.clear();
                for (uint32_t kk = rowptr[r0]; kk < rowptr[r1]; ++kk) cols.push_back(rowadj[kk] & 0xffffu);
                std::sort(cols.begin(), cols.end());
                cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
                for (uint32_t c : cols) tccol[at++] = (uint16_t)c;
                for (uint32_t kk = rowptr[r0]; kk < rowptr[r1]; ++kk)
                    tcslot[kk] = (uint16_t)(std::lower_bound(cols.begin(), cols.end(), rowadj[kk] & 0xffffu) - cols.begin());
                max_tc = std::max<uint32_t>(max_tc, (uint32_t)cols.size());
            }
            tcptr[s.n_tile] = at;
            h->max_tc = max_tc;
        }
        uint32_t k = 0;
        for (uint32_t t = 0; t <= s.n_tile; ++t) {
            while (k < s.nnz && (nz[k].i >> 3) < t) ++k;
            tileelem[t] = k;
        }
    }
    // CSR over tile rows
    {
        uint32_t oi = 0;
        for (uint32_t t = 0; t <= s.n_tile; ++t) {
            while (oi < s.n_octile && oct[oi].trow < t) ++oi;
            tilerow[t] = oi;
        }
    }
    if (g->pool_bytes) std::memcpy(base + s.off_pool, g->pool, g->pool_bytes);
    return GDB_OK;
}

// ---------------------------------------------------------------------------
// batch packing: k graphs from column-concatenated inputs on a few host threads
// ---------------------------------------------------------------------------
extern "C" int gdb_graphs_pack_batch(const gdb_layout *L, const gdb_batch_src *src, uint64_t *blob_off, void *blobs,
                                     uint64_t capacity, int32_t n_threads) {
    if (!L || !src || !blob_off || !src->node_off || !src->edge_off)
        return gdb_fail(GDB_ERR_INVALID, "gdb_graphs_pack_batch: null argument");
    const uint32_t n = src->n_graphs;
    auto graph_at = [&](uint32_t g) {
        gdb_graph_src s{};
        const uint64_t n0 = src->node_off[g], e0 = src->edge_off[g];
        s.n_node = (uint32_t)(src->node_off[g + 1] - n0);
        s.n_edge = (uint32_t)(src->edge_off[g + 1] - e0);
        s.nodes = static_cast<const uint8_t *>(src->nodes) + n0 * L->node_size;
        s.edge_i = src->edge_i + e0;
        s.edge_j = src->edge_j + e0;
        s.edge_w = src->edge_w ? src->edge_w + e0 : nullptr;
        s.edge_labels = src->edge_labels ? static_cast<const uint8_t *>(src->edge_labels) + e0 * L->edge_label_size : nullptr;
        if (src->pool_off) {
            s.pool = static_cast<const uint8_t *>(src->pool) + src->pool_off[g];
            s.pool_bytes = (uint32_t)(src->pool_off[g + 1] - src->pool_off[g]);
        }
        return s;
    };
    unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt, std::max(1u, n / 64u)));
    std::vector<int> status(nt, GDB_OK);
    std::vector<std::string> message(nt);
    auto run = [&](auto &&body) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                const uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
                for (uint32_t g = lo; g < hi && status[t] == GDB_OK; ++g) {
                    const int rc = body(g);
                    if (rc != GDB_OK) {
                        status[t] = rc;
                        message[t] = "graph " + std::to_string(g) + ": " + gdb_last_error();
                    }
                }
            });
        for (auto &t : th) t.join();
        for (unsigned t = 0; t < nt; ++t)
            if (status[t] != GDB_OK) return gdb_fail(status[t], "%s", message[t].c_str());
        return GDB_OK;
    };
    if (!blobs) {
        int rc = run([&](uint32_t g) {
            const gdb_graph_src s = graph_at(g);
            uint64_t bytes = 0;
            const int r = gdb_graph_packed_size(L, &s, &bytes);
            blob_off[g + 1] = bytes;
            return r;
        });
        if (rc) return rc;
        blob_off[0] = 0;
        for (uint32_t g = 0; g < n; ++g) blob_off[g + 1] += blob_off[g];
        return GDB_OK;
    }
    if (blob_off[n] > capacity) return gdb_fail(GDB_ERR_INVALID, "gdb_graphs_pack_batch: blob buffer too small");
    return run([&](uint32_t g) {
        const gdb_graph_src s = graph_at(g);
        return gdb_graph_pack(L, &s, static_cast<uint8_t *>(blobs) + blob_off[g], blob_off[g + 1] - blob_off[g]);
    });
}

// ---------------------------------------------------------------------------
// node reordering (host): permutations for Graph.permute that shrink the tile
// footprint of a graph.  Replaces reference graphdot/graph/reorder/rcm.py:7-22
// (scipy) and plays the role of reference graphdot/graph/reorder/pbr (a
// hypergraph partitioner around kahypar that minimises non-empty 8 x 8 tiles,
// pbr/mnom.py:11-24) with a native greedy tile-growing partition.
// ---------------------------------------------------------------------------
namespace {

struct Csr {
    std::vector<uint32_t> ptr, adj;
};

bool build_csr(uint32_t n, uint32_t m, const uint32_t *ei, const uint32_t *ej, Csr &g) {
    g.ptr.assign(n + 1, 0);
    for (uint32_t k = 0; k < m; ++k) {
        if (ei[k] >= n || ej[k] >= n) return false;
        if (ei[k] == ej[k]) continue;
        g.ptr[ei[k] + 1]++;
        g.ptr[ej[k] + 1]++;
    }
    for (uint32_t i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
    g.adj.resize(g.ptr[n]);
    std::vector<uint32_t> fill(g.ptr.begin(), g.ptr.end() - 1);
    for (uint32_t k = 0; k < m; ++k) {
        if (ei[k] == ej[k]) continue;
        g.adj[fill[ei[k]]++] = ej[k];
        g.adj[fill[ej[k]]++] = ei[k];
    }
    return true;
}

// reverse Cuthill-McKee: BFS from a low-degree node of every component, neighbours by
// increasing degree, order reversed
void order_rcm(uint32_t n, const Csr &g, uint32_t *perm) {
    auto deg = [&](uint32_t v) { return g.ptr[v + 1] - g.ptr[v]; };
    std::vector<uint32_t> by_degree(n), order;
    std::vector<char> seen(n, 0);
    for (uint32_t i = 0; i < n; ++i) by_degree[i] = i;
    std::stable_sort(by_degree.begin(), by_degree.end(), [&](uint32_t a, uint32_t b) { return deg(a) < deg(b); });
    order.reserve(n);
    std::vector<uint32_t> nb;
    for (uint32_t s : by_degree) {
        if (seen[s]) continue;
        seen[s] = 1;
        size_t head = order.size();
        order.push_back(s);
        while (head < order.size()) {
            const uint32_t v = order[head++];
            nb.clear();
            for (uint32_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k)
                if (!seen[g.adj[k]]) {
                    seen[g.adj[k]] = 1;
                    nb.push_back(g.adj[k]);
                }
            std::stable_sort(nb.begin(), nb.end(), [&](uint32_t a, uint32_t b) { return deg(a) < deg(b); });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    for (uint32_t k = 0; k < n; ++k) perm[k] = order[n - 1 - k];
}

// greedy tile growing: fill one block of 8 nodes at a time with the unassigned node that
// has the most neighbours in the block (ties: in the previous block, then the smaller
// degree); a new block is seeded next to the previous one.  Neighbourhoods end up inside
// a tile row / the adjacent one, i.e. few non-empty 8 x 8 tiles.
void order_tiles(uint32_t n, const Csr &g, uint32_t *perm) {
    auto deg = [&](uint32_t v) { return g.ptr[v + 1] - g.ptr[v]; };
    std::vector<int> in_cur(n, 0), in_prev(n, 0);  // neighbours in the current / previous block
    std::vector<char> done(n, 0);
    std::vector<uint32_t> touched, prev_block, block;
    uint32_t placed = 0;
    while (placed < n) {
        block.clear();
        for (int slot = 0; slot < 8 && placed < n; ++slot) {
            // candidates: unassigned nodes adjacent to the current or the previous block
            long best = -1;
            long best_key = -1;
            for (uint32_t v : touched) {
                if (done[v]) continue;
                const long key = (long)in_cur[v] * 4096 * 64 + (long)in_prev[v] * 4096 + (4095 - (long)std::min<uint32_t>(deg(v), 4095));
                if (key > best_key) {
                    best_key = key;
                    best = v;
                }
            }
            if (best < 0) {  // nothing adjacent: the unassigned node of smallest degree
                uint32_t bd = ~0u;
                for (uint32_t v = 0; v < n; ++v)
                    if (!done[v] && deg(v) < bd) {
                        bd = deg(v);
                        best = v;
                    }
            }
            const uint32_t v = (uint32_t)best;
            done[v] = 1;
            perm[placed++] = v;
            block.push_back(v);
            for (uint32_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k) {
                const uint32_t u = g.adj[k];
                if (!in_cur[u] && !in_prev[u]) touched.push_back(u);
                in_cur[u]++;
            }
        }
        // the finished block becomes the previous one
        for (uint32_t v : prev_block)
            for (uint32_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k) in_prev[g.adj[k]]--;
        for (uint32_t v : block)
            for (uint32_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k) {
                in_cur[g.adj[k]]--;
                in_prev[g.adj[k]]++;
            }
        prev_block = block;
        std::vector<uint32_t> keep;
        for (uint32_t u : touched)
            if (!done[u] && (in_cur[u] || in_prev[u])) keep.push_back(u);
        std::sort(keep.begin(), keep.end());
        keep.erase(std::unique(keep.begin(), keep.end()), keep.end());
        touched.swap(keep);
    }
}

}  // namespace

extern "C" int gdb_graph_reorder(uint32_t n_node, uint32_t n_edge, const uint32_t *edge_i, const uint32_t *edge_j,
                                 int32_t method, uint32_t *perm) {
    if (!n_node || !perm || (n_edge && (!edge_i || !edge_j))) return gdb_fail(GDB_ERR_INVALID, "gdb_graph_reorder: null argument");
    Csr g;
    if (!build_csr(n_node, n_edge, edge_i, edge_j, g)) return gdb_fail(GDB_ERR_INVALID, "edge end point out of range");
    if (method == GDB_REORDER_RCM)
        order_rcm(n_node, g, perm);
    else if (method == GDB_REORDER_TILES)
        order_tiles(n_node, g, perm);
    else
        return gdb_fail(GDB_ERR_INVALID, "unknown reordering method %d", method);
    return GDB_OK;
}

extern "C" int gdb_graph_count_tiles(uint32_t n_node, uint32_t n_edge, const uint32_t *edge_i, const uint32_t *edge_j,
                                     const uint32_t *perm, uint64_t *n_tiles) {
    if (!n_tiles || (n_edge && (!edge_i || !edge_j))) return gdb_fail(GDB_ERR_INVALID, "gdb_graph_count_tiles: null argument");
    std::vector<uint32_t> inv;
    if (perm) {
        inv.resize(n_node);
        for (uint32_t k = 0; k < n_node; ++k) {
            if (perm[k] >= n_node) return gdb_fail(GDB_ERR_INVALID, "perm is not a permutation");
            inv[perm[k]] = k;  // new index of old node perm[k] is k (Graph.permute)
        }
    }
    std::vector<uint64_t> keys;
    keys.reserve(2 * (size_t)n_edge);
    for (uint32_t k = 0; k < n_edge; ++k) {
        if (edge_i[k] >= n_node || edge_j[k] >= n_node) return gdb_fail(GDB_ERR_INVALID, "edge end point out of range");
        const uint64_t i = perm ? inv[edge_i[k]] : edge_i[k], j = perm ? inv[edge_j[k]] : edge_j[k];
        keys.push_back((i >> 3) << 32 | (j >> 3));
        keys.push_back((j >> 3) << 32 | (i >> 3));
    }
    std::sort(keys.begin(), keys.end());
    *n_tiles = (uint64_t)(std::unique(keys.begin(), keys.end()) - keys.begin());
    return GDB_OK;
}
