// mlgk_large.cuh -- solver for LARGE graph pairs (BASELINE configuration C4:
// 200-500 nodes per graph, N = n1 n2 = 4e4 ... 2.5e5 product-graph nodes): one
// thread-block CLUSTER per pair.  Same algorithm and results contract as
// mlgk_solve (mlgk_solver.cuh); replaces reference
// graphdot/cpp/marginalized_kernel.h:189-490 (compute), :492-804
// (compute_duo), :806-997 (derivative) and the job loop of reference
// kernel/marginalized/template.cu:29-475 for pairs in this regime.
//
// Why a cluster: the CG vectors of one pair are 5 N floats = 0.8 ... 5 MB and
// one matvec is 2e6 gathered products; GDB_CLUSTER CTAs split both, which cuts
// the latency of a pair and the number of arenas in flight.  (Measured, DESIGN.md
// section 4.3: the kernel is bound by instruction issue and latency, not by
// HBM or L2 bandwidth -- two resident CTAs per SM matter more than L2 residency.)
//
// Per pair (G1 = rows, G2 = columns, element i = i1 n2p + i2 with the row
// stride n2p = n2 rounded up to 4):
//  * the tile rows (8 rows) of G1 are dealt to the CTAs of the cluster in
//    contiguous blocks; a CTA owns the elements of its rows in ALL vector passes;
//  * matvec of one tile row: the rows of the search direction p that the tile
//    row's elements touch -- the packer's sorted list of distinct neighbour
//    columns of that tile row (gdb_pack.cpp: tcptr / tccol / tcslot), about 14
//    rows for a banded graph instead of the 24 rows of its 3 octiles -- are
//    staged in shared memory with cp.async (16-byte LDGSTS, L2 -> smem) into ONE
//    buffer: the second resident CTA of the SM (another pair) covers the
//    staging latency, and the shared memory a second buffer would take is
//    worth more as L1 (32.5 k vs 42.9 k pairs/s on C4).  The CTA's warps
//    share out (rows of the tile row) x (blocks of 32 columns); a warp's lanes
//    own 32 consecutive columns at a time: every gather of p is a
//    conflict-free LDS.  Rows of 1..8 elements run a gather loop specialised
//    on the element count, with the row's elements in registers;
//  * G2 is held per CTA in shared memory in ELL form (slot-major:
//    ell[t][position] = {neighbour, edge}, positions = columns sorted by
//    decreasing degree so that the lanes of a warp run the same number of
//    slots), so that the lanes' loads are conflict-free as well; the elements of the tile row of G1 are staged next
//    to the rows they gather from and read by broadcast loads;
//    the edge microkernel is evaluated on the fly (nnz1 nnz2 = 2e6 products
//    per matvec do not fit on chip);
//  * dot products: warp shuffles -> shared memory -> one DSMEM store per CTA
//    pair (st.shared::cluster) -> barrier.cluster; every CTA sums the same
//    partials in the same order, so alpha / beta are bit-identical cluster-wide;
//  * x, r, p, A p, diag (and y) live in this cluster's slice of a global arena
//    and are streamed with 128-bit loads; pad columns hold zeros.
// No atomics, fixed summation order => bit-reproducible results.
//
// Macros from the generated header: GDB_CLUSTER (CTAs per pair), GDB_LELL (ELL
// slots per column held in shared memory; further neighbours are read from
// global memory).
#pragma once

#ifndef GDB_CLUSTER
#define GDB_CLUSTER 2
#endif
#ifndef GDB_LELL
#define GDB_LELL 12
#endif
#ifndef GDB_LBLOCK
// 8 warps share out the (rows of a staging step) x (blocks of 32 columns).  Two resident CTAs
// of 256 threads leave ptxas 128 registers per thread: no spills in the gather loops, and 16
// warps per SM then do what 32 warps at the 64-register cap of 512 threads do with 316 B of
// spill stores / 596 B of spill loads per thread.  All 500 C4 graphs: 256 threads 46.8 k pairs/s,
// 512 threads 43.5 k; with the Jacobian 22.4 k / 22.0 k.  (When a badly ordered graph set leaves
// room for ONE CTA per SM only, 512 threads are better, 29.1 k vs 26.7 k: reorder the graphs.)
#define GDB_LBLOCK 256
#endif
#ifndef GDB_LTR
#define GDB_LTR 1        // tile rows of G1 per staging step
#endif

#if GDB_NODAL == 0 && (GDB_BUILD_MASK & 4)  // graph-level outputs only; nodal outputs run in mlgk_solve

__device__ __forceinline__ unsigned gdb_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned gdb_cluster_id() {
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void gdb_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the same shared-memory variable of CTA `rank` of this cluster (DSMEM)
__device__ __forceinline__ void gdb_st_cluster(float *local, unsigned rank, float v) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(gdb_smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ void gdb_st_cluster_u64(unsigned long long *local, unsigned rank, unsigned long long v) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(gdb_smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(ra), "l"(v) : "memory");
}
__device__ __forceinline__ void gdb_cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void gdb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void gdb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct gdb_large_shared {
    float warp[GDB_LBLOCK / 32][4];     // per-warp partial sums
    float part[2][GDB_CLUSTER][4];      // per-CTA partial sums, written by every CTA of the cluster
    unsigned long long job;             // next job, written by the cluster's first CTA
};

// Sum of K <= 4 values over the whole cluster; one block barrier + one cluster
// barrier.  Every CTA returns the same bits.
template<int K> __device__ __forceinline__ void gdb_cluster_sum(float (&v)[K], gdb_large_shared &S, int &flip) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float t = gdb_warp_sum(v[k]);
        if (lane == 0) S.warp[warp][k] = t;
    }
    __syncthreads();
    if (threadIdx.x < GDB_CLUSTER * K) {
        const unsigned k = threadIdx.x % K, dst = threadIdx.x / K;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < GDB_LBLOCK / 32; ++w) t += S.warp[w][k];
        gdb_st_cluster(&S.part[flip][gdb_cluster_rank()][k], dst, t);
    }
    gdb_cluster_sync();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < GDB_CLUSTER; ++c) t += S.part[flip][c][k];
        v[k] = t;
    }
    flip ^= 1;
}

struct gdb_large_graph {
    const float *degree;
    const node_t *node;
    const edge_t *edge;
    const unsigned *rowptr, *rowadj, *rowpos, *tcptr, *lanemap;
    const unsigned short *tccol, *tcslot;
    int n, nnz, n_tile, max_degree, max_tc;
};

__device__ __forceinline__ gdb_large_graph gdb_large_view(const unsigned char *base) {
    const gdb_graph_hdr *h = reinterpret_cast<const gdb_graph_hdr *>(base);
    gdb_large_graph v;
    v.degree = reinterpret_cast<const float *>(base + h->off_degree);
    v.node = reinterpret_cast<const node_t *>(base + h->off_node);
    v.edge = reinterpret_cast<const edge_t *>(base + h->off_edge);
    v.rowptr = reinterpret_cast<const unsigned *>(base + h->off_rowptr);
    v.rowadj = reinterpret_cast<const unsigned *>(base + h->off_rowadj);
    v.rowpos = reinterpret_cast<const unsigned *>(base + h->off_ellslot);
    v.tcptr = reinterpret_cast<const unsigned *>(base + h->off_tcptr);
    v.lanemap = reinterpret_cast<const unsigned *>(base + h->off_lanemap);
    v.tccol = reinterpret_cast<const unsigned short *>(base + h->off_tccol);
    v.tcslot = reinterpret_cast<const unsigned short *>(base + h->off_tcslot);
    v.n = h->n_node;
    v.nnz = h->nnz;
    v.n_tile = h->n_tile;
    v.max_degree = (int)h->max_degree;
    v.max_tc = (int)h->max_tc;
    return v;
}

// One ELL entry of G2 in shared memory: byte offset of the neighbour inside a
// staged row and the edge -- one 64-bit load for 4-byte edge types.
struct __align__(8) gdb_ell_t {
    unsigned off;
    edge_t e;
};
static_assert(sizeof(gdb_ell_t) % 8 == 0, "ELL entries are loaded 8 bytes at a time");
extern "C" __device__ const unsigned gdb_large_layout[3] = {(unsigned)sizeof(gdb_ell_t), GDB_LBLOCK, GDB_LTR};

// load an entry through its 32-bit shared-window address (always LDS, no generic-address arithmetic)
__device__ __forceinline__ gdb_ell_t gdb_lds_ell(unsigned addr) {
    unsigned w[sizeof(gdb_ell_t) / 4];
#pragma unroll
    for (unsigned k = 0; k < sizeof(gdb_ell_t) / 8; ++k)
        asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w[2 * k]), "=r"(w[2 * k + 1]) : "r"(addr + 8u * k));
    gdb_ell_t out;
    memcpy(&out, w, sizeof out);
    return out;
}

// Everything a CTA needs to sweep its tile rows of the product graph.
struct gdb_large_ctx {
    gdb_large_graph g1, g2;
    unsigned n2p;                 // row stride of the vectors (floats)
    int t_lo, t_hi;               // this CTA's tile rows of G1
    gdb_ell_t *ell;               // [D2][n2p] slot t of the column at position p of G2
    unsigned *colinfo;            // [n2p] position p (columns sorted by decreasing degree) -> column | degree << 16
    int D2;                       // ELL slots in shared memory
    float *stage[2];              // staged rows of the gathered vector
    gdb_ell_t *rowel;             // elements of ALL of this CTA's rows of G1 (CSR order), filled once per pair:
                                  // {byte offset, inside a staging buffer, of the staged row the element
                                  // gathers from; edge}
    unsigned k_lo;                // CSR position of the first of them
    unsigned cap;                 // elements held in shared memory
    bool dbl;                     // both staging buffers usable
};

// cp.async the rows of `vec` that the tile rows [t, t_end) of G1 touch into staging
// buffer b (one list of rows per tile row, back to back)
__device__ __forceinline__ void gdb_large_stage(const gdb_large_ctx &C, const float *vec, int t, int t_end, int b) {
    const unsigned per_row = C.n2p / 4u;  // 16-byte chunks per row
    const unsigned dst0 = gdb_smem_u32(C.stage[b]);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned c0 = C.g1.tcptr[t], cnt = C.g1.tcptr[t_end] - c0;
    // a warp copies whole rows, its lanes consecutive 16-byte chunks (no index division)
    for (unsigned s = threadIdx.x >> 5; s < cnt; s += GDB_LBLOCK / 32) {
        const float *src = vec + (size_t)C.g1.tccol[c0 + s] * C.n2p;
        const unsigned dst = dst0 + s * C.n2p * 4u;
#pragma unroll 4
        for (unsigned ch = lane; ch < per_row; ch += 32u) gdb_cp_async16(dst + ch * 16u, src + ch * 4u);
    }
}

// acc[] += (edge value | edge Jacobian) of (e1, e2) times the staged vector entry
template<int MODE, int NACC> __device__ __forceinline__ void gdb_large_product(const gdb_params &P, const edge_t &e1,
                                                                               const edge_t &e2, float pj, float (&acc)[NACC]) {
    if constexpr (MODE == 0) {
        acc[0] = fmaf(gdb_edge_value(P, e1, e2), pj, acc[0]);
    } else {
#if GDB_NE > 0
        float de[GDB_NE];
        P.edge_kernel.jacobian(e1.label, e2.label, de);
#if GDB_WEIGHTED
        pj *= e1.weight * e2.weight;
#endif
#pragma unroll
        for (int a = 0; a < GDB_NE; ++a) acc[a] = fmaf(de[a], pj, acc[a]);
#endif
    }
}

template<int V> struct gdb_int {
    static constexpr int value = V;
};

// One sweep over this CTA's tile rows with `vec` staged.
//  MODE 0 (matvec):  out[i] = diag[i] vec[i] - sum_j W_ij vec[j]; returns sum vec[i] out[i] in res[0]
//  MODE 1 (edge Jacobian): res[m] += sum_i yv[i] sum_j w1 w2 dkE_m(e1, e2) vec[j]
template<int MODE> __device__ __forceinline__ void gdb_large_sweep(const gdb_params &P, const gdb_large_ctx &C,
                                                                   const float *vec, const float *__restrict__ diag,
                                                                   float *__restrict__ out, const float *__restrict__ yv,
                                                                   float *res) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int n1 = C.g1.n, n2 = C.g2.n;
    const unsigned n2p = C.n2p;
    // a step covers GDB_LTR tile rows: fewer barriers, and 8 GDB_LTR rows x column blocks
    // deal out evenly over the 32 warps
    if (C.t_lo < C.t_hi) {
        gdb_large_stage(C, vec, C.t_lo, min(C.t_lo + GDB_LTR, C.t_hi), 0);
        gdb_cp_async_commit();
    }
    for (int t = C.t_lo, step = 0; t < C.t_hi; t += GDB_LTR, ++step) {
        const int b = C.dbl ? (step & 1) : 0;
        const int t_end = min(t + GDB_LTR, C.t_hi);
        if (C.dbl && t_end < C.t_hi) {
            gdb_large_stage(C, vec, t_end, min(t_end + GDB_LTR, C.t_hi), b ^ 1);
            gdb_cp_async_commit();
            gdb_cp_async_wait<1>();
        } else {
            gdb_cp_async_wait<0>();
        }
        __syncthreads();  // staged rows of this step visible to every warp
        // warps = (rows of this step) x (groups of column blocks): a warp keeps ONE row -- its
        // elements stay in registers across all of the warp's column blocks
        const unsigned rows_here = (unsigned)(min(8 * t_end, n1) - 8 * t);
        // groups of column blocks per row so that rows x groups deals out evenly over the
        // warps: warps / gcd(rows, warps) (= warps / rows when that divides, e.g. 16 / 8)
        unsigned groups;
        {
            constexpr unsigned NW = GDB_LBLOCK / 32;
            unsigned a = rows_here, b = NW;
            while (b) {
                const unsigned r = a % b;
                a = b;
                b = r;
            }
            groups = NW / a;
        }
        const unsigned ell_sa = gdb_smem_u32(C.ell), ell_stride = n2p * (unsigned)sizeof(gdb_ell_t);
        const unsigned stage_sa = gdb_smem_u32(C.stage[b]);
        constexpr int NACC = MODE == 0 ? 1 : (GDB_NE > 0 ? GDB_NE : 1);
        double dres[MODE == 0 ? 1 : NACC];  // MODE 1: partial sums of the step in double (see the epilogue)
#pragma unroll
        for (int a = 0; a < (MODE == 0 ? 1 : NACC); ++a) dres[a] = 0.0;
#pragma unroll 1
        for (unsigned wi = warp; wi < rows_here * groups; wi += GDB_LBLOCK / 32) {
            const unsigned q = wi / rows_here;
            const int i1 = 8 * t + (int)(wi - q * rows_here);
            const unsigned k1beg = C.g1.rowptr[i1], deg1 = C.g1.rowptr[i1 + 1] - k1beg;
            // elements of this row that sit in shared memory (the rest, rare, in global memory)
            const unsigned u_sh = k1beg - C.k_lo >= C.cap ? 0u : min(deg1, C.cap - (k1beg - C.k_lo));
            const unsigned row_sa = gdb_smem_u32(C.rowel) + (k1beg - C.k_lo) * (unsigned)sizeof(gdb_ell_t);
            // K > 0: the row has exactly K elements, held in registers and fully unrolled
            // (about 8 instructions per product: sub, 2 mul, ex2, fma for a square-exponential
            // edge kernel + 2 LDS + 1 add, instead of 17 with a rolled loop over a handful of
            // elements); K = 0: any other count (and the Jacobian sweep, once per pair)
            auto row_pass = [&](auto kc) {
                constexpr int K = decltype(kc)::value;
                gdb_ell_t el[K > 0 ? K : 1];
#pragma unroll
                for (int u = 0; u < K; ++u) {
                    el[u] = gdb_lds_ell(row_sa + (unsigned)u * (unsigned)sizeof(gdb_ell_t));  // broadcast loads
                    el[u].off += stage_sa;  // offset inside a staging buffer -> shared-window address
                }
#pragma unroll 1
                for (unsigned c0 = 32u * q; c0 < (unsigned)n2; c0 += 32u * groups) {
                    // lanes take 32 consecutive POSITIONS of the degree-sorted column order (the
                    // packer's lane map, stable within a degree): the lanes of a warp then have
                    // (almost) the same number of slots -- in node order the per-lane trip count
                    // of the slot loop left 28 % of the lanes idle
                    const unsigned pos = c0 + lane;
                    const bool live = pos < (unsigned)n2;
                    const unsigned ci = live ? C.colinfo[pos] : 0u;
                    const unsigned c = ci & 0xffffu;
                    // own element: its global loads are issued before the gather loop, which
                    // hides their latency
                    const size_t i_own = (size_t)i1 * n2p + (live ? c : 0u);
                    float own_v, own_d = 0.f;
                    if constexpr (MODE == 0) {
                        own_v = vec[i_own];
                        own_d = diag[i_own];
                    } else {
                        own_v = yv[i_own];
                    }
                    const unsigned d2 = ci >> 16;
                    const unsigned d2s = min(d2, (unsigned)C.D2);
                    unsigned at = ell_sa + pos * (unsigned)sizeof(gdb_ell_t);
                    float acc[NACC];
#pragma unroll
                    for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll 1
                    for (unsigned t2 = 0; t2 < d2s; ++t2, at += ell_stride) {  // slots of this lane's column
                        const gdb_ell_t en = gdb_lds_ell(at);
                        if constexpr (K > 0) {
#pragma unroll
                            for (int u = 0; u < K; ++u) {
                                float pj;
                                asm("ld.shared.f32 %0, [%1];" : "=f"(pj) : "r"(el[u].off + en.off));
                                gdb_large_product<MODE, NACC>(P, el[u].e, en.e, pj, acc);
                            }
                        } else {
                            unsigned ra = row_sa;
#pragma unroll 2
                            for (unsigned u = 0; u < u_sh; ++u, ra += (unsigned)sizeof(gdb_ell_t)) {  // warp-uniform
                                const gdb_ell_t e1 = gdb_lds_ell(ra);
                                float pj;
                                asm("ld.shared.f32 %0, [%1];" : "=f"(pj) : "r"(stage_sa + e1.off + en.off));
                                gdb_large_product<MODE, NACC>(P, e1.e, en.e, pj, acc);
                            }
                        }
                    }
                    if (u_sh < deg1 || d2 > d2s) {  // rare: elements beyond the shared-memory copies
                        const unsigned kb = C.g2.rowptr[c];
                        for (unsigned u = 0; u < deg1; ++u) {
                            const unsigned k1 = k1beg + u;
                            const edge_t e1 = C.g1.edge[C.g1.rowadj[k1] >> 16];
                            const float *row = C.stage[b] + (C.g1.tcptr[i1 >> 3] - C.g1.tcptr[t] + (unsigned)C.g1.tcslot[k1]) * n2p;
                            for (unsigned t2 = (u < u_sh ? d2s : 0u); t2 < d2; ++t2) {
                                const unsigned a2 = C.g2.rowadj[kb + t2];
                                gdb_large_product<MODE, NACC>(P, e1, C.g2.edge[a2 >> 16], row[a2 & 0xffffu], acc);
                            }
                        }
                    }
                    if (live) {  // own element
                        if constexpr (MODE == 0) {
                            const float r = fmaf(own_d, own_v, -acc[0]);
                            out[i_own] = r;
                            res[0] = fmaf(own_v, r, res[0]);
                        } else {
#pragma unroll
                            for (int a = 0; a < NACC; ++a) dres[a] += (double)(own_v * acc[a]);
                        }
                    }
                }
            };
            switch (MODE == 0 && u_sh == deg1 ? u_sh : 0u) {
            case 1: row_pass(gdb_int<1>{}); break;
            case 2: row_pass(gdb_int<2>{}); break;
            case 3: row_pass(gdb_int<3>{}); break;
            case 4: row_pass(gdb_int<4>{}); break;
            case 5: row_pass(gdb_int<5>{}); break;
            case 6: row_pass(gdb_int<6>{}); break;
            case 7: row_pass(gdb_int<7>{}); break;
            case 8: row_pass(gdb_int<8>{}); break;
            default: row_pass(gdb_int<0>{}); break;
            }
        }
        if constexpr (MODE == 1) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) res[a] += (float)dres[a];
        }
        __syncthreads();  // every warp is done with buffer b before it is refilled
        if (!C.dbl && t_end < C.t_hi) {
            gdb_large_stage(C, vec, t_end, min(t_end + GDB_LTR, C.t_hi), 0);
            gdb_cp_async_commit();
        }
    }
}

// Jacobi-PCG over the cluster for A x = rhs, x0 = 0.  On entry r = rhs on this
// CTA's elements [e_lo, e_hi); on exit x holds the solution there.
__device__ __forceinline__ int gdb_large_pcg(const gdb_params &P, const gdb_large_ctx &C, const float *__restrict__ diag,
                                             float *__restrict__ x, float *__restrict__ r, float *p, float *__restrict__ Ap,
                                             size_t e_lo, size_t e_hi, int N, float tol, gdb_large_shared &S, int &flip) {
    float4 *x4 = reinterpret_cast<float4 *>(x), *r4 = reinterpret_cast<float4 *>(r), *p4 = reinterpret_cast<float4 *>(p);
    const float4 *d4 = reinterpret_cast<const float4 *>(diag), *a4 = reinterpret_cast<const float4 *>(Ap);
    const size_t q_lo = e_lo / 4, q_hi = e_hi / 4;  // float4 range (rows are multiples of 4 floats)
    float s1[1] = {0.f};
    for (size_t i = q_lo + threadIdx.x; i < q_hi; i += GDB_LBLOCK) {
        const float4 rv = r4[i], dv = d4[i];
        const float4 z = make_float4(__fdividef(rv.x, dv.x), __fdividef(rv.y, dv.y), __fdividef(rv.z, dv.z), __fdividef(rv.w, dv.w));
        x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        p4[i] = z;
        s1[0] += rv.x * z.x + rv.y * z.y + rv.z * z.z + rv.w * z.w;
    }
    gdb_cluster_sum(s1, S, flip);  // cluster barrier: p is complete everywhere
    float rho = s1[0];
    const float thresh2 = (tol * (float)N) * (tol * (float)N);
    int k = 0;
    while (k < N && rho != 0.f) {
        float pAp[1] = {0.f};
        gdb_large_sweep<0>(P, C, p, diag, Ap, nullptr, pAp);
        gdb_cluster_sum(pAp, S, flip);
        if (pAp[0] == 0.f) break;
        ++k;
        const float alpha = __fdividef(rho, pAp[0]);
        float s2[2] = {0.f, 0.f};
        // (two or four float4 per thread and vector in flight -- loads batched ahead of the
        // arithmetic -- were slower: the kernel sits at its 128-register cap, the batch spills
        // into the gather loops; 44.0 k -> 36.2 k / 33.1 k pairs/s on 100 C4 graphs -- also
        // with the sweep kept out of line, which by itself costs 10 %)
        for (size_t i = q_lo + threadIdx.x; i < q_hi; i += GDB_LBLOCK) {
            const float4 pv = p4[i], av = a4[i], dv = d4[i];
            float4 xv = x4[i], rv = r4[i];
            xv.x = fmaf(alpha, pv.x, xv.x), xv.y = fmaf(alpha, pv.y, xv.y), xv.z = fmaf(alpha, pv.z, xv.z), xv.w = fmaf(alpha, pv.w, xv.w);
            rv.x = fmaf(-alpha, av.x, rv.x), rv.y = fmaf(-alpha, av.y, rv.y), rv.z = fmaf(-alpha, av.z, rv.z), rv.w = fmaf(-alpha, av.w, rv.w);
            x4[i] = xv;
            r4[i] = rv;
            s2[0] += rv.x * rv.x + rv.y * rv.y + rv.z * rv.z + rv.w * rv.w;
            s2[1] += rv.x * __fdividef(rv.x, dv.x) + rv.y * __fdividef(rv.y, dv.y) + rv.z * __fdividef(rv.z, dv.z) +
                     rv.w * __fdividef(rv.w, dv.w);
        }
        gdb_cluster_sum(s2, S, flip);
        if (s2[0] < thresh2) break;
        const float beta = __fdividef(s2[1], rho);
        for (size_t i = q_lo + threadIdx.x; i < q_hi; i += GDB_LBLOCK) {
            const float4 pv = p4[i], rv = r4[i], dv = d4[i];
            p4[i] = make_float4(fmaf(beta, pv.x, __fdividef(rv.x, dv.x)), fmaf(beta, pv.y, __fdividef(rv.y, dv.y)),
                                fmaf(beta, pv.z, __fdividef(rv.z, dv.z)), fmaf(beta, pv.w, __fdividef(rv.w, dv.w)));
        }
        rho = s2[1];
        gdb_cluster_sync();  // p complete everywhere before the next matvec stages its rows
    }
    return k;
}

#ifndef GDB_LMINB
#define GDB_LMINB 2  // resident CTAs per SM asked of ptxas: two CTAs (of different pairs) per SM overlap
                     // each other's latency-bound phases (vector passes, barriers) with gather loops
#endif
extern "C" __global__ void __cluster_dims__(GDB_CLUSTER, 1, 1) __launch_bounds__(GDB_LBLOCK, GDB_LMINB)
    mlgk_solve_large(const __grid_constant__ gdb_params P) {
    extern __shared__ __align__(16) unsigned char gdb_smem[];
    __shared__ gdb_large_shared S;
    int flip = 0;
    const gdb_params_fixed &F = P.f;
    const unsigned rank = gdb_cluster_rank();
#if GDB_GRADIENT
    constexpr int NVEC = 6;
#else
    constexpr int NVEC = 5;
#endif
    float *const arena = F.scratch + (size_t)gdb_cluster_id() * F.scratch_stride;
    gdb_cluster_sync();  // every CTA of the cluster has started: its shared memory may be written remotely

    while (true) {
        // ---- the cluster's first CTA claims a job and tells the others (DSMEM) ------
        if (rank == 0 && threadIdx.x < GDB_CLUSTER) {
            unsigned long long job = 0;
            if (threadIdx.x == 0) job = atomicAdd(F.counters, 1ull);
            job = __shfl_sync((1u << GDB_CLUSTER) - 1u, job, 0);
            gdb_st_cluster_u64(&S.job, threadIdx.x, job);
        }
        gdb_cluster_sync();
        const unsigned long long job = S.job;
        if (job >= F.n_jobs) break;
        unsigned ja, jb;
        gdb_decode_job(F, job, ja, jb);
        const gdb_graph_ref ref1 = F.graphs[ja], ref2 = F.graphs[jb];
        gdb_large_ctx C;
        C.g1 = gdb_large_view(ref1.blob);
        C.g2 = gdb_large_view(ref2.blob);
        const int n1 = C.g1.n, n2 = C.g2.n, N = n1 * n2;
        const unsigned n2p = ((unsigned)n2 + 3u) & ~3u;
        C.n2p = n2p;
        // tile rows of G1 dealt to the CTAs in contiguous, balanced blocks
        C.t_lo = (int)((long long)C.g1.n_tile * rank / GDB_CLUSTER);
        C.t_hi = (int)((long long)C.g1.n_tile * (rank + 1) / GDB_CLUSTER);
        const int row_lo = min(8 * C.t_lo, n1), row_hi = min(8 * C.t_hi, n1);
        const size_t e_lo = (size_t)row_lo * n2p, e_hi = (size_t)row_hi * n2p;
        const size_t Npad = (size_t)n1 * n2p;

        // ---- shared memory: ELL copy of G2 | staging buffers ---------------------------
        C.D2 = min(C.g2.max_degree, GDB_LELL);
        C.cap = F.pad3;
        unsigned off = 0;
        C.ell = reinterpret_cast<gdb_ell_t *>(gdb_smem);
        off += (((unsigned)C.D2 * n2p * (unsigned)sizeof(gdb_ell_t)) + 15u) & ~15u;
        C.colinfo = reinterpret_cast<unsigned *>(gdb_smem + off);
        off += ((n2p * 4u) + 15u) & ~15u;
        C.rowel = reinterpret_cast<gdb_ell_t *>(gdb_smem + off);
        off += ((C.cap * (unsigned)sizeof(gdb_ell_t)) + 15u) & ~15u;
        const unsigned buf_bytes = (unsigned)GDB_LTR * (unsigned)C.g1.max_tc * n2p * 4u;
        C.stage[0] = reinterpret_cast<float *>(gdb_smem + off);
        C.dbl = off + 2u * buf_bytes <= F.smem_bytes;
        C.stage[1] = C.dbl ? reinterpret_cast<float *>(gdb_smem + off + buf_bytes) : C.stage[0];
        // the elements of this CTA's rows of G1, once per pair (the staging steps of every
        // matvec then issue cp.async only)
        C.k_lo = C.g1.rowptr[row_lo];
        for (unsigned k = threadIdx.x; k < min(C.g1.rowptr[row_hi] - C.k_lo, C.cap); k += GDB_LBLOCK) {
            const unsigned kk = C.k_lo + k;
            const unsigned tile = (C.g1.rowpos[kk] & 0xffffu) >> 3;  // tile row of this element
            const unsigned first = (unsigned)C.t_lo + (tile - (unsigned)C.t_lo) / GDB_LTR * GDB_LTR;  // first tile row of its step
            gdb_ell_t el;
            el.off = (C.g1.tcptr[tile] - C.g1.tcptr[first] + (unsigned)C.g1.tcslot[kk]) * n2p * 4u;
            el.e = C.g1.edge[C.g1.rowadj[kk] >> 16];
            C.rowel[k] = el;
        }
        for (unsigned ps = threadIdx.x; ps < n2p; ps += GDB_LBLOCK) {
            unsigned ci = 0u;
            if (ps < (unsigned)n2) {
                const unsigned c = C.g2.lanemap[ps] & 0xffffu;
                ci = c | ((C.g2.rowptr[c + 1] - C.g2.rowptr[c]) << 16);
            }
            C.colinfo[ps] = ci;
        }
        for (unsigned k2 = threadIdx.x; k2 < (unsigned)C.g2.nnz; k2 += GDB_LBLOCK) {
            const unsigned rp = C.g2.rowpos[k2], c = rp & 0xffffu, t2 = rp >> 16;
            if (t2 < (unsigned)C.D2) {
                const unsigned a2 = C.g2.rowadj[k2];
                gdb_ell_t en;
                en.off = (a2 & 0xffffu) * 4u;
                en.e = C.g2.edge[a2 >> 16];
                C.ell[t2 * n2p + (C.g2.lanemap[c] >> 16)] = en;
            }
        }

        float *x = arena, *r = arena + Npad, *p = arena + 2 * Npad, *Ap = arena + 3 * Npad, *diag = arena + 4 * Npad;
        const float Q = 1.0f / (1.0f - F.q), Q2 = Q * Q;

        // ---- setup on this CTA's rows: diag = Dx / Vx, rhs = Dx; pad columns neutral ----
        for (size_t e = e_lo + threadIdx.x; e < e_hi; e += GDB_LBLOCK) {
            const unsigned i1 = (unsigned)(e / n2p), i2 = (unsigned)(e - (size_t)i1 * n2p);
            float d = 1.f, b = 0.f;
            if (i2 < (unsigned)n2) {
                b = C.g1.degree[i1] * C.g2.degree[i2] * Q2;
                d = __fdividef(b, P.node_kernel(C.g1.node[i1], C.g2.node[i2]));
            }
            diag[e] = d;
            r[e] = b;
            Ap[e] = 0.f;
        }
        __syncthreads();  // ELL copy complete (the first cluster barrier inside the solve orders the rest)
        int iters = gdb_large_pcg(P, C, diag, x, r, p, Ap, e_lo, e_hi, N, F.ftol, S, flip);
#if GDB_GRADIENT
        float *y = arena + 5 * Npad;
        for (size_t e = e_lo + threadIdx.x; e < e_hi; e += GDB_LBLOCK) {
            const unsigned i1 = (unsigned)(e / n2p), i2 = (unsigned)(e - (size_t)i1 * n2p);
            r[e] = i2 < (unsigned)n2 ? P.p_start(C.g1.node[i1]) * P.p_start(C.g2.node[i2]) : 0.f;
        }
        __syncthreads();  // the solve reads r four elements at a time: other threads' writes
        iters += gdb_large_pcg(P, C, diag, y, r, p, Ap, e_lo, e_hi, N, F.ftol, S, flip);
#endif
        if (rank == 0 && threadIdx.x == 0) {
            atomicAdd(F.counters + 1, (unsigned long long)iters);
            atomicAdd(F.counters + 2, (unsigned long long)iters * (unsigned long long)C.g1.nnz * (unsigned long long)C.g2.nnz);
            atomicAdd(F.counters + 3, (unsigned long long)iters * (unsigned long long)N);
        }

        // ---- epilogue over this CTA's elements --------------------------------------------
        constexpr int NACC = 1 + (GDB_GRADIENT ? GDB_NP + 1 + GDB_NV : 0);
        // per-thread partial sums in double: a thread adds ~N / 512 terms of one sign and similar
        // size; a float accumulator rounds them the same way for long stretches (label-free
        // kernels at N = 9e5: K off by 1.2e-5 relative with float partial sums)
        double dacc[NACC];
#pragma unroll
        for (int m = 0; m < NACC; ++m) dacc[m] = 0.0;
        for (size_t e = e_lo + threadIdx.x; e < e_hi; e += GDB_LBLOCK) {
            const unsigned i1 = (unsigned)(e / n2p), i2 = (unsigned)(e - (size_t)i1 * n2p);
            if (i2 >= (unsigned)n2) continue;
            const node_t &u1 = C.g1.node[i1];
            const node_t &u2 = C.g2.node[i2];
            const float p1 = P.p_start(u1), p2 = P.p_start(u2);
            const float dx = C.g1.degree[i1] * C.g2.degree[i2] * Q2;
            const float v = __fdividef(dx, diag[e]);  // Vx recovered from the cached diagonal
            const float xi = x[e];
            float xs = xi;
#if GDB_LMIN == 1
            xs -= v;
#endif
            dacc[0] += (double)(xs * (p1 * p2));
#if GDB_GRADIENT
            const float yi = y[e];
#if GDB_NP > 0
            {
                float d1[GDB_NP], d2[GDB_NP];
                P.p_start.jacobian(u1, d1);
                P.p_start.jacobian(u2, d2);
#pragma unroll
                for (int m = 0; m < GDB_NP; ++m) dacc[1 + m] += (double)(fmaf(d1[m], p2, p1 * d2[m]) * xs);
            }
#endif
            dacc[1 + GDB_NP] += (double)(2.f * Q * dx * yi * (1.f - __fdividef(xi, v)));
#if GDB_NV > 0
            {
                float dv[GDB_NV];
                P.node_kernel.jacobian(u1, u2, dv);
                const float cc = yi * xi * __fdividef(dx, v * v);
#pragma unroll
                for (int m = 0; m < GDB_NV; ++m) {
                    float t = cc * dv[m];
#if GDB_LMIN == 1
                    t -= p1 * p2 * dv[m];
#endif
                    dacc[2 + GDB_NP + m] += (double)t;
                }
            }
#endif
#endif
        }
        float acc[NACC];
#pragma unroll
        for (int m = 0; m < NACC; ++m) acc[m] = (float)dacc[m];
#if GDB_GRADIENT && GDB_NE > 0
        float eacc[GDB_NE];
#pragma unroll
        for (int m = 0; m < GDB_NE; ++m) eacc[m] = 0.f;
        gdb_large_sweep<1>(P, C, x, nullptr, nullptr, y, eacc);
#endif
#pragma unroll
        for (int m0 = 0; m0 < NACC; m0 += 4) {
            float part[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) part[k] = (m0 + k < NACC) ? acc[m0 + k] : 0.f;
            gdb_cluster_sum(part, S, flip);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m0 + k < NACC) acc[m0 + k] = part[k];
        }
#if GDB_GRADIENT && GDB_NE > 0
#pragma unroll
        for (int m0 = 0; m0 < GDB_NE; m0 += 4) {
            float part[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) part[k] = (m0 + k < GDB_NE) ? eacc[m0 + k] : 0.f;
            gdb_cluster_sum(part, S, flip);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m0 + k < GDB_NE) eacc[m0 + k] = part[k];
        }
#endif
        if (rank == 0 && threadIdx.x == 0) {
            const unsigned I1 = F.starts[ja] - F.row0, I2 = F.starts[jb] - F.col0;
            const unsigned long long plane = (unsigned long long)F.nX * F.nY;
            (void)plane;
            (void)I2;
#if !GDB_DIAGONAL
            const float norm_rs = gdb_norm_scale(F, ja, jb);
            acc[0] *= norm_rs;
#endif
#if GDB_DIAGONAL
            F.gram[I1] = acc[0];
#else
            F.gram[(unsigned long long)I1 + (unsigned long long)I2 * F.nX] = acc[0];
#if GDB_SYMMETRIC
            if (ja != jb) F.gram[(unsigned long long)I2 + (unsigned long long)I1 * F.nX] = acc[0];
#endif
#endif
#if GDB_GRADIENT
#pragma unroll
            for (int m = 0; m < GDB_NJ; ++m) {
                float val;
                if (m < GDB_NP + 1 + GDB_NV) {
                    val = acc[1 + m];
                } else {
#if GDB_NE > 0
                    val = eacc[m - (GDB_NP + 1 + GDB_NV)];
#else
                    val = 0.f;
#endif
                }
#if GDB_DIAGONAL
                F.grad[(unsigned long long)I1 + (unsigned long long)m * F.nX] = val;
#else
                val = gdb_norm_grad(F, ja, jb, m, norm_rs, acc[0], val);
                F.grad[(unsigned long long)I1 + (unsigned long long)I2 * F.nX + m * plane] = val;
#if GDB_SYMMETRIC
                if (ja != jb) F.grad[(unsigned long long)I2 + (unsigned long long)I1 * F.nX + m * plane] = val;
#endif
#endif
            }
#endif
        }
        // the next job's setup overwrites the arena and the shared-memory tables: every
        // CTA of the cluster must be done reading them (the job-claim barrier at the top
        // of the loop provides that)
    }
}

#endif  // GDB_NODAL == 0
