// Randomised driver for the pure-host entry points of the library (octile packer, batch
// packer, node reorderings), built by tests/test_native_sanitized.py with
// -fsanitize=address,undefined: memory safety and UB on ragged inputs -- single nodes, no
// edges, isolated nodes, self loops, parallel edges, sizes around the 8- and 32-boundaries,
// weighted and unweighted, with and without a variable-length feature pool.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "graphdot_b200.h"

static thread_local char g_err[512];
// stand-ins for the error sink of gdb_abi.cpp (which needs the CUDA runtime)
extern "C" const char *gdb_last_error(void) { return g_err; }
int gdb_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            fprintf(stderr, "%s:%d: %s failed (%s)\n", __FILE__, __LINE__, #cond, g_err); \
            exit(1);                                                             \
        }                                                                        \
    } while (0)

struct node_t {
    float x;
    uint64_t feat_data;  // frozen_array<float>: pool-relative offset, relocated by the packer
    int32_t feat_size;
    int32_t pad;
};
struct label_t {
    float length;
};

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 300;
    std::mt19937 rng(12345);
    gdb_layout L{};
    L.node_size = sizeof(node_t);
    L.edge_label_size = sizeof(label_t);
    L.edge_label_align = alignof(label_t);
    L.n_node_ptr = 1;
    L.node_ptr_offset[0] = offsetof(node_t, feat_data);
    uint64_t total_bytes = 0;
    const uint32_t sizes[] = {1, 2, 7, 8, 9, 31, 32, 33, 63, 64, 65, 100, 257, 600};
    for (int round = 0; round < rounds; ++round) {
        const uint32_t n = sizes[rng() % (sizeof sizes / sizeof *sizes)];
        const int kind = rng() % 4;  // 0 no edges, 1 sparse, 2 dense-ish, 3 with self loops + duplicates
        L.weighted = (int32_t)(rng() % 2);
        uint32_t m = kind == 0 ? 0 : kind == 1 ? n : std::min<uint32_t>(4 * n, n * (n - 1) / 2 + 3);
        std::vector<uint32_t> ei, ej;
        std::vector<float> ew;
        std::vector<label_t> lab;
        for (uint32_t k = 0; k < m; ++k) {
            uint32_t a = rng() % n, b = rng() % n;
            if (kind != 3 && a == b) continue;
            ei.push_back(a), ej.push_back(b);
            if (kind == 3 && (rng() % 4) == 0) ei.push_back(a), ej.push_back(b);  // parallel edge
        }
        m = (uint32_t)ei.size();
        for (uint32_t k = 0; k < m; ++k) ew.push_back(0.5f + (rng() % 100) * 0.01f), lab.push_back({(rng() % 100) * 0.01f});
        std::vector<node_t> nodes(n);
        std::vector<float> pool;
        for (uint32_t i = 0; i < n; ++i) {
            const int len = 1 + (int)(rng() % 5);
            nodes[i] = {(float)i, (uint64_t)(pool.size() * sizeof(float)), len, 0};
            for (int t = 0; t < len; ++t) pool.push_back((float)t);
        }
        gdb_graph_src src{};
        src.n_node = n, src.n_edge = m, src.nodes = nodes.data();
        src.edge_i = ei.data(), src.edge_j = ej.data();
        src.edge_w = L.weighted ? ew.data() : nullptr;
        src.edge_labels = lab.data();
        src.pool = pool.data(), src.pool_bytes = (uint32_t)(pool.size() * sizeof(float));
        uint64_t bytes = 0;
        CHECK(gdb_graph_packed_size(&L, &src, &bytes) == GDB_OK);
        CHECK(bytes >= 96 && bytes % 16 == 0);
        std::vector<unsigned char> blob(bytes);
        CHECK(gdb_graph_pack(&L, &src, blob.data(), bytes) == GDB_OK);
        CHECK(gdb_graph_pack(&L, &src, blob.data(), bytes - 16) != GDB_OK);  // too small: refused
        int32_t hdr[4];
        memcpy(hdr, blob.data(), sizeof hdr);
        CHECK(hdr[0] == (int32_t)n);
        total_bytes += bytes;

        // reorderings: permutations, tile count consistent with relabelled edges
        for (int method : {GDB_REORDER_RCM, GDB_REORDER_TILES}) {
            std::vector<uint32_t> perm(n, ~0u), seen(n, 0);
            CHECK(gdb_graph_reorder(n, m, ei.data(), ej.data(), method, perm.data()) == GDB_OK);
            for (uint32_t k = 0; k < n; ++k) {
                CHECK(perm[k] < n && !seen[perm[k]]);
                seen[perm[k]] = 1;
            }
            uint64_t t_perm = 0, t_relabelled = 0;
            CHECK(gdb_graph_count_tiles(n, m, ei.data(), ej.data(), perm.data(), &t_perm) == GDB_OK);
            std::vector<uint32_t> inv(n), ri(m), rj(m);
            for (uint32_t k = 0; k < n; ++k) inv[perm[k]] = k;
            for (uint32_t k = 0; k < m; ++k) ri[k] = inv[ei[k]], rj[k] = inv[ej[k]];
            CHECK(gdb_graph_count_tiles(n, m, ri.data(), rj.data(), nullptr, &t_relabelled) == GDB_OK);
            CHECK(t_perm == t_relabelled);
        }
        if (m) {  // an end point out of range is refused everywhere
            std::vector<uint32_t> bad(ei);
            bad[rng() % m] = n;
            gdb_graph_src s2 = src;
            s2.edge_i = bad.data();
            uint64_t b2;
            std::vector<uint32_t> perm(n);
            CHECK(gdb_graph_packed_size(&L, &s2, &b2) != GDB_OK);
            CHECK(gdb_graph_reorder(n, m, bad.data(), ej.data(), 0, perm.data()) != GDB_OK);
            CHECK(gdb_graph_count_tiles(n, m, bad.data(), ej.data(), nullptr, &b2) != GDB_OK);
        }
    }

    // batch packer == per-graph packer (no pool: scalar attributes only)
    {
        gdb_layout B{};
        B.node_size = 8, B.edge_label_size = 4, B.edge_label_align = 4, B.weighted = 0;
        const uint32_t k = 37;
        std::vector<uint64_t> noff(k + 1, 0), eoff(k + 1, 0);
        std::vector<uint64_t> nodes;
        std::vector<uint32_t> ei, ej;
        std::vector<float> lab;
        for (uint32_t g = 0; g < k; ++g) {
            const uint32_t n = 1 + rng() % 40, m = rng() % (3 * n);
            for (uint32_t i = 0; i < n; ++i) nodes.push_back(rng());
            for (uint32_t e = 0; e < m; ++e) {
                const uint32_t a = rng() % n, b = rng() % n;
                if (a == b) continue;
                ei.push_back(a), ej.push_back(b), lab.push_back((float)(rng() % 10));
            }
            noff[g + 1] = nodes.size(), eoff[g + 1] = ei.size();
        }
        gdb_batch_src bs{};
        bs.n_graphs = k, bs.node_off = noff.data(), bs.edge_off = eoff.data();
        bs.nodes = nodes.data(), bs.edge_i = ei.data(), bs.edge_j = ej.data(), bs.edge_labels = lab.data();
        std::vector<uint64_t> boff(k + 1, 0);
        for (int threads : {1, 3, 0}) {
            CHECK(gdb_graphs_pack_batch(&B, &bs, boff.data(), nullptr, 0, threads) == GDB_OK);
            std::vector<unsigned char> blobs(boff[k]);
            CHECK(gdb_graphs_pack_batch(&B, &bs, boff.data(), blobs.data(), blobs.size(), threads) == GDB_OK);
            for (uint32_t g = 0; g < k; ++g) {
                gdb_graph_src s{};
                s.n_node = (uint32_t)(noff[g + 1] - noff[g]), s.n_edge = (uint32_t)(eoff[g + 1] - eoff[g]);
                s.nodes = nodes.data() + noff[g];
                s.edge_i = ei.data() + eoff[g], s.edge_j = ej.data() + eoff[g], s.edge_labels = lab.data() + eoff[g];
                uint64_t bytes = 0;
                CHECK(gdb_graph_packed_size(&B, &s, &bytes) == GDB_OK);
                CHECK(bytes == boff[g + 1] - boff[g]);
                std::vector<unsigned char> one(bytes);
                CHECK(gdb_graph_pack(&B, &s, one.data(), bytes) == GDB_OK);
                CHECK(memcmp(one.data(), blobs.data() + boff[g], bytes) == 0);
            }
        }
    }
    printf("pack_fuzz ok: %d graphs, %llu blob bytes\n", rounds, (unsigned long long)total_bytes);
    return 0;
}
