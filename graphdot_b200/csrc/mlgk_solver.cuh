// mlgk_solver.cuh -- the fixed hand-written sm_100a solver the microkernel
// expressions are spliced into (NVRTC translation unit = mlgk_prelude.cuh +
// generated types/functors/macros + this file).
//
// One persistent group of GDB_BLOCK threads (a CTA; a single warp for small
// pairs) pulls graph-pair jobs from a global counter and, per pair (G1, G2),
// solves the product-graph system of the marginalized graph kernel
//     (diag(Dx / Vx) - Ax o Ex) x = Dx            
// with Jacobi-preconditioned CG, then reduces x with the starting
// probabilities into the Gram entry and, on request, evaluates the adjoint
// Jacobian with respect to every hyper-parameter.
//
// Replaces reference graphdot/kernel/marginalized/template.cu:29-475
// (job loop, epilogues) and graphdot/cpp/marginalized_kernel.h:164-997
// (load / compute / compute_duo / derivative).  Differences by design:
//   * the matvec is a gather: each thread owns output elements (i1, i2) and
//     walks the octiles of tile-row i1/8 of G1 and i2/8 of G2 through their
//     row bit masks -- no atomics, bit-reproducible results (the reference
//     scatters with float atomicAdd, marginalized_kernel.h:326-346);
//   * Vx and Dx are evaluated once per pair into a cached diagonal (the
//     reference re-evaluates the node kernel twice per iteration, :227, :421);
//   * graphs are position-independent blobs staged whole into shared memory;
//     all CG vectors live in shared memory when the pair fits, else in a
//     per-CTA global arena;
//   * dot products are warp-shuffle trees + one shared-memory hop (the
//     reference uses shared-memory float atomics, util_cuda.h:33-49).
//
// Macros supplied by the generated header:
//   GDB_BLOCK, GDB_MIN_BLOCKS, GDB_WEIGHTED, GDB_DIAGONAL, GDB_SYMMETRIC,
//   GDB_NODAL (0 graph-level, 1 nodal, 2 block), GDB_LMIN, GDB_GRADIENT,
//   GDB_NP, GDB_NV, GDB_NE (Jacobian sizes); types node_t, edge_label_t,
//   node_kernel_t, edge_kernel_t, p_start_t.
#pragma once

#ifndef GDB_MIN_BLOCKS
#define GDB_MIN_BLOCKS 1
#endif

#define GDB_NJ (GDB_NP + 1 + GDB_NV + GDB_NE)
#define GDB_WARPS (GDB_BLOCK / 32)

#if GDB_WEIGHTED
struct edge_t {
    float32 weight;
    edge_label_t label;
};
#else
struct edge_t {
    edge_label_t label;
};
#endif

// ---- packed graph blob (host mirror: csrc/gdb_pack.cpp) --------------------
struct gdb_graph_hdr {
    int n_node, n_octile, nnz, n_tile;
    unsigned off_degree, off_node, off_octile, off_tilerow;
    unsigned off_edge, off_pool, blob_bytes, flags;
    unsigned off_emeta, off_rowptr, off_rowadj, off_tileelem;
    unsigned max_degree, off_ellslot, off_lanemap, vcols;
    unsigned off_tcptr, off_tccol, off_tcslot, max_tc;
};

struct gdb_octile {
    unsigned long long mask;  // bit (8 * row + col)
    unsigned start;           // index of the first element in edge[]
    unsigned short trow, tcol;
};

struct gdb_graph_ref {
    const unsigned char *blob;
    unsigned bytes, n_node;
};

struct gdb_params_fixed {
    const gdb_graph_ref *graphs;
    const uint2 *jobs;
    const unsigned *starts;
    float *gram;
    float *grad;
    float *scratch;
    unsigned long long *counters;  // [0] next job, [1] CG iterations, [2] products, [3] vector elements
    unsigned long long scratch_stride;
    unsigned long long n_jobs;
    unsigned job_mode, i0, i1, j0, j1, nX, nY, nJ;
    float q, eps, ftol, gtol;
    unsigned smem_bytes, row0, col0, norm_n;
    // fused normalization K_ij / sqrt(K_ii K_jj) (reference kernel/fix.py:46-73):
    // self-similarities and their Jacobians per graph, or null
    const float *norm_diag;
    const float *norm_ddiag;  // [m * norm_n + graph]
    unsigned blob_slot;       // small kernel: bytes of one blob staging buffer (two graphs)
    unsigned pad3;            // large kernel: elements of a tile row held in shared memory
};

struct gdb_params {
    gdb_params_fixed f;
    alignas(16) node_kernel_t node_kernel;
    alignas(16) edge_kernel_t edge_kernel;
    alignas(16) p_start_t p_start;
};

// Offsets follow from the alignas(16) members above (NVRTC has no offsetof).
__host__ __device__ constexpr unsigned gdb_up16(unsigned v) { return (v + 15u) & ~15u; }
#define GDB_OFF_V gdb_up16((unsigned)sizeof(gdb_params_fixed))
#define GDB_OFF_E gdb_up16(GDB_OFF_V + (unsigned)sizeof(node_kernel_t))
#define GDB_OFF_P gdb_up16(GDB_OFF_E + (unsigned)sizeof(edge_kernel_t))
static_assert(alignof(node_kernel_t) <= 16 && alignof(edge_kernel_t) <= 16 && alignof(p_start_t) <= 16,
              "hyper-parameter structs must not be over-aligned");
static_assert(sizeof(gdb_params) == gdb_up16(GDB_OFF_P + (unsigned)sizeof(p_start_t)), "parameter block layout");

// which kernels this translation unit holds: 1 mlgk_solve, 2 mlgk_solve_small, 4 mlgk_solve_large.
// The host library compiles one NVRTC module per kernel, concurrently (gdb_abi.cpp).
#ifndef GDB_BUILD_MASK
#define GDB_BUILD_MASK 7
#endif
extern "C" __device__ const unsigned gdb_param_layout[8] = {
    (unsigned)sizeof(gdb_params), GDB_OFF_V, GDB_OFF_E, GDB_OFF_P,
    (unsigned)sizeof(gdb_params_fixed), (unsigned)sizeof(node_t), (unsigned)sizeof(edge_t), (unsigned)GDB_NJ};

struct gdb_graph_view {
    const float *degree;
    const node_t *node;
    const gdb_octile *oct;
    const unsigned *trow;
    const edge_t *edge;
    const unsigned *rowptr, *rowadj;  // row index: (col | element << 16) per row
    int n, n_octile, nnz;
    bool index16;                     // row index valid (n, nnz < 65536)
};

__device__ __forceinline__ gdb_graph_view gdb_view(const unsigned char *base) {
    const gdb_graph_hdr *h = reinterpret_cast<const gdb_graph_hdr *>(base);
    gdb_graph_view v;
    v.degree = reinterpret_cast<const float *>(base + h->off_degree);
    v.node = reinterpret_cast<const node_t *>(base + h->off_node);
    v.oct = reinterpret_cast<const gdb_octile *>(base + h->off_octile);
    v.trow = reinterpret_cast<const unsigned *>(base + h->off_tilerow);
    v.edge = reinterpret_cast<const edge_t *>(base + h->off_edge);
    v.rowptr = reinterpret_cast<const unsigned *>(base + h->off_rowptr);
    v.rowadj = reinterpret_cast<const unsigned *>(base + h->off_rowadj);
    v.n = h->n_node;
    v.n_octile = h->n_octile;
    v.nnz = h->nnz;
    v.index16 = (h->flags & 2u) != 0;
    return v;
}

// Normalization of a graph-level Gram entry and of its Jacobian, applied in
// the epilogue when the self-similarities are available on the device.
__device__ __forceinline__ float gdb_norm_scale(const gdb_params_fixed &F, unsigned ja, unsigned jb) {
    return F.norm_diag ? 1.0f / sqrtf(F.norm_diag[ja] * F.norm_diag[jb]) : 1.0f;
}
__device__ __forceinline__ float gdb_norm_grad(const gdb_params_fixed &F, unsigned ja, unsigned jb, int m, float rs,
                                               float kn, float raw) {
    if (!F.norm_diag) return raw;
    const float la = __fdividef(F.norm_ddiag[(unsigned long long)m * F.norm_n + ja], F.norm_diag[ja]);
    const float lb = __fdividef(F.norm_ddiag[(unsigned long long)m * F.norm_n + jb], F.norm_diag[jb]);
    return fmaf(raw, rs, -0.5f * kn * (la + lb));
}

__device__ __forceinline__ float gdb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void gdb_group_sync() {
#if GDB_BLOCK == 32
    __syncwarp();
#else
    __syncthreads();
#endif
}

// Sum over the whole group; `red` is a [2][GDB_WARPS] shared buffer and
// `flip` alternates between its halves so that one barrier per reduction is
// enough.  Every thread returns the same value (fixed summation order).
__device__ __forceinline__ float gdb_group_sum(float v, float *red, int &flip) {
    v = gdb_warp_sum(v);
#if GDB_BLOCK == 32
    return v;
#else
    float *buf = red + flip * GDB_WARPS;
    flip ^= 1;
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < GDB_WARPS; ++w) t += buf[w];
    return t;
#endif
}

__device__ __forceinline__ float gdb_edge_value(const gdb_params &P, const edge_t &a, const edge_t &b) {
#if GDB_WEIGHTED
    return a.weight * b.weight * P.edge_kernel(a.label, b.label);
#else
    return P.edge_kernel(a.label, b.label);
#endif
}

// out[i] = diag[i] * in[i] - sum_j W[i, j] in[j] for the elements this thread
// owns (i = tid, tid + GDB_BLOCK, ...); returns the partial sum of in[i]*out[i].
__device__ __forceinline__ float gdb_matvec(const gdb_params &P, const gdb_graph_view &g1,
                                            const gdb_graph_view &g2, const float *__restrict__ diag,
                                            const float *in, float *__restrict__ out) {
    const int n2 = g2.n, N = g1.n * g2.n;
    float dot = 0.f;
    for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
        const int i1 = i / n2, i2 = i - i1 * n2;
        const int r1 = (i1 & 7) * 8, r2 = (i2 & 7) * 8;
        const unsigned o1_end = g1.trow[(i1 >> 3) + 1], o2_beg = g2.trow[i2 >> 3], o2_end = g2.trow[(i2 >> 3) + 1];
        float acc = 0.f;
        for (unsigned o1 = g1.trow[i1 >> 3]; o1 < o1_end; ++o1) {
            const gdb_octile t1 = g1.oct[o1];
            unsigned m1 = (unsigned)(t1.mask >> r1) & 0xffu;
            if (!m1) continue;
            const edge_t *e1 = g1.edge + t1.start + __popcll(t1.mask & ((1ull << r1) - 1ull));
            const float *in1 = in + (int)t1.tcol * 8 * n2;
            for (; m1; m1 &= m1 - 1, ++e1) {
                const float *inrow = in1 + (__ffs(m1) - 1) * n2;
                for (unsigned o2 = o2_beg; o2 < o2_end; ++o2) {
                    const gdb_octile t2 = g2.oct[o2];
                    unsigned m2 = (unsigned)(t2.mask >> r2) & 0xffu;
                    if (!m2) continue;
                    const edge_t *e2 = g2.edge + t2.start + __popcll(t2.mask & ((1ull << r2) - 1ull));
                    const float *inp = inrow + (int)t2.tcol * 8;
                    for (; m2; m2 &= m2 - 1, ++e2) {
                        acc = fmaf(gdb_edge_value(P, *e1, *e2), inp[__ffs(m2) - 1], acc);
                    }
                }
            }
        }
        const float v = in[i];
        const float r = fmaf(diag[i], v, -acc);
        out[i] = r;
        dot = fmaf(v, r, dot);
    }
    return dot;
}

// Same product through the row index (CSR x CSR): no bit-mask decoding, two
// nested loops over the neighbours of i1 and i2; the edge microkernel is
// evaluated on the fly (the cached-W variant lives in mlgk_small.cuh).
__device__ __forceinline__ float gdb_matvec_csr(const gdb_params &P, const gdb_graph_view &g1,
                                                const gdb_graph_view &g2, const float *__restrict__ diag,
                                                const float *__restrict__ in, float *__restrict__ out) {
    const unsigned n2 = (unsigned)g2.n, N = (unsigned)(g1.n * g2.n);
    const float inv_n2 = __frcp_rn((float)n2);
    float dot = 0.f;
    for (unsigned i = threadIdx.x; i < N; i += GDB_BLOCK) {
        unsigned i1 = (unsigned)(__uint2float_rz(i) * inv_n2);
        unsigned i2 = i - i1 * n2;
        if ((int)i2 < 0) {
            --i1;
            i2 += n2;
        } else if (i2 >= n2) {
            ++i1;
            i2 -= n2;
        }
        const unsigned k1end = g1.rowptr[i1 + 1], k2beg = g2.rowptr[i2], k2end = g2.rowptr[i2 + 1];
        float acc = 0.f;
        for (unsigned k1 = g1.rowptr[i1]; k1 < k1end; ++k1) {
            const unsigned a1 = g1.rowadj[k1];
            const edge_t &e1 = g1.edge[a1 >> 16];
            const float *inrow = in + (a1 & 0xffffu) * n2;
            for (unsigned k2 = k2beg; k2 < k2end; ++k2) {
                const unsigned a2 = g2.rowadj[k2];
                acc = fmaf(gdb_edge_value(P, e1, g2.edge[a2 >> 16]), inrow[a2 & 0xffffu], acc);
            }
        }
        const float v = in[i];
        const float r = fmaf(diag[i], v, -acc);
        out[i] = r;
        dot = fmaf(v, r, dot);
    }
    return dot;
}

// Sum of two values over the group with one barrier.
__device__ __forceinline__ void gdb_group_sum2(float &a, float &b, float *red, int &flip) {
    a = gdb_warp_sum(a);
    b = gdb_warp_sum(b);
#if GDB_BLOCK > 32
    float *buf = red + flip * GDB_WARPS;  // red holds 2 x 2 x GDB_WARPS floats
    flip ^= 1;
    if ((threadIdx.x & 31) == 0) {
        buf[threadIdx.x >> 5] = a;
        buf[2 * GDB_WARPS + (threadIdx.x >> 5)] = b;
    }
    __syncthreads();
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < GDB_WARPS; ++w) {
        ta += buf[w];
        tb += buf[2 * GDB_WARPS + w];
    }
    a = ta;
    b = tb;
#endif
}

// Jacobi-PCG for A x = rhs with x0 = 0.  On entry r = rhs; on exit x holds the
// solution; r, p, Ap are clobbered.  Returns the number of iterations.  The
// vectors do not alias (__restrict__) and the streaming loops are unrolled by
// four with all loads issued before the first store, so that a thread keeps
// several independent global/L2 requests in flight (large pairs stream their
// vectors from L2/HBM).
__device__ __forceinline__ int gdb_pcg(const gdb_params &P, const gdb_graph_view &g1, const gdb_graph_view &g2,
                                       const float *__restrict__ diag, float *__restrict__ x, float *__restrict__ r,
                                       float *__restrict__ p, float *__restrict__ Ap, float tol, float *red, int &flip) {
    const int N = g1.n * g2.n;
    float rho = 0.f;
#pragma unroll 4
    for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
        const float ri = r[i];
        const float z = __fdividef(ri, diag[i]);
        x[i] = 0.f;
        p[i] = z;
        rho = fmaf(ri, z, rho);
    }
    rho = gdb_group_sum(rho, red, flip);
    gdb_group_sync();  // p complete before the first matvec
    const float thresh = tol * (float)N;
    int k = 0;
    while (k < N && rho != 0.f) {
        float pAp = (g1.index16 && g2.index16) ? gdb_matvec_csr(P, g1, g2, diag, p, Ap) : gdb_matvec(P, g1, g2, diag, p, Ap);
        pAp = gdb_group_sum(pAp, red, flip);
        if (pAp == 0.f) break;
        ++k;
        const float alpha = __fdividef(rho, pAp);
        float rr = 0.f, rz = 0.f;
        int i = threadIdx.x;
        for (; i + 3 * GDB_BLOCK < N; i += 4 * GDB_BLOCK) {
            float pv[4], av[4], xv[4], rv[4], dv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                pv[u] = p[i + u * GDB_BLOCK];
                av[u] = Ap[i + u * GDB_BLOCK];
                xv[u] = x[i + u * GDB_BLOCK];
                rv[u] = r[i + u * GDB_BLOCK];
                dv[u] = diag[i + u * GDB_BLOCK];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                x[i + u * GDB_BLOCK] = fmaf(alpha, pv[u], xv[u]);
                const float ri = fmaf(-alpha, av[u], rv[u]);
                r[i + u * GDB_BLOCK] = ri;
                rr = fmaf(ri, ri, rr);
                rz = fmaf(ri, __fdividef(ri, dv[u]), rz);
            }
        }
        for (; i < N; i += GDB_BLOCK) {
            x[i] = fmaf(alpha, p[i], x[i]);
            const float ri = fmaf(-alpha, Ap[i], r[i]);
            r[i] = ri;
            rr = fmaf(ri, ri, rr);
            rz = fmaf(ri, __fdividef(ri, diag[i]), rz);
        }
        gdb_group_sum2(rr, rz, red, flip);
        if (sqrtf(rr) < thresh) break;
        const float beta = __fdividef(rz, rho);
        i = threadIdx.x;
        for (; i + 3 * GDB_BLOCK < N; i += 4 * GDB_BLOCK) {
            float pv[4], rv[4], dv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                pv[u] = p[i + u * GDB_BLOCK];
                rv[u] = r[i + u * GDB_BLOCK];
                dv[u] = diag[i + u * GDB_BLOCK];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) p[i + u * GDB_BLOCK] = fmaf(beta, pv[u], __fdividef(rv[u], dv[u]));
        }
        for (; i < N; i += GDB_BLOCK) p[i] = fmaf(beta, p[i], __fdividef(r[i], diag[i]));
        rho = rz;
        gdb_group_sync();  // p complete before the next matvec
    }
    return k;
}

__device__ __forceinline__ void gdb_decode_job(const gdb_params_fixed &f, unsigned long long idx, unsigned &a,
                                               unsigned &b) {
    if (f.job_mode == 0) {
        const uint2 j = f.jobs[idx];
        a = j.x;
        b = j.y;
    } else if (f.job_mode == 1) {
        const unsigned long long nj = f.j1 - f.j0;
        a = f.i0 + (unsigned)(idx / nj);
        b = f.j0 + (unsigned)(idx % nj);
    } else {
        // rows i in [i0, i1), columns j in [i, j1): local row r holds m - r entries
        const long long m = (long long)f.j1 - (long long)f.i0;
        const double md = (double)m;
        long long row = (long long)floor(((2.0 * md + 1.0) - sqrt((2.0 * md + 1.0) * (2.0 * md + 1.0) - 8.0 * (double)idx)) * 0.5);
        if (row < 0) row = 0;
        while (row > 0 && (unsigned long long)(row * m - row * (row - 1) / 2) > idx) --row;
        while ((unsigned long long)((row + 1) * m - (row + 1) * row / 2) <= idx) ++row;
        const unsigned long long first = (unsigned long long)(row * m - row * (row - 1) / 2);
        a = f.i0 + (unsigned)row;
        b = a + (unsigned)(idx - first);
    }
}

__device__ __forceinline__ void gdb_copy16(void *dst, const void *src, unsigned bytes) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    for (unsigned k = threadIdx.x; k < bytes / 16; k += GDB_BLOCK) d[k] = s[k];
}

#if GDB_BUILD_MASK & 1
extern "C" __global__ void __launch_bounds__(GDB_BLOCK, GDB_MIN_BLOCKS)
    mlgk_solve(const __grid_constant__ gdb_params P) {
    extern __shared__ __align__(16) unsigned char gdb_smem[];
    __shared__ unsigned long long s_job;
    __shared__ float s_red[4 * (GDB_WARPS > 0 ? GDB_WARPS : 1)];
    int flip = 0;
    const gdb_params_fixed &F = P.f;
#if GDB_GRADIENT
    constexpr int NVEC = 6;
#else
    constexpr int NVEC = 5;
#endif

    while (true) {
        gdb_group_sync();  // previous job's shared memory is dead
        if (threadIdx.x == 0) s_job = atomicAdd(F.counters, 1ull);
        gdb_group_sync();
        const unsigned long long job = s_job;
        if (job >= F.n_jobs) break;
        unsigned ja, jb;
        gdb_decode_job(F, job, ja, jb);
        const gdb_graph_ref ref1 = F.graphs[ja], ref2 = F.graphs[jb];
        const unsigned N = ref1.n_node * ref2.n_node;

        // ---- placement ---------------------------------------------------------
        // The row index and the edge elements of both graphs (what every matvec
        // reads) are staged in shared memory when they fit; the rest of the blob
        // (nodes, feature pools, octiles) is read from global memory.  The CG
        // vectors live in shared memory when they fit as well, else in this
        // CTA's slice of the global arena.
        const bool same = (ja == jb);
        gdb_graph_view g1 = gdb_view(ref1.blob), g2 = gdb_view(ref2.blob);
        unsigned used = 0;
        {
            auto idx_bytes = [](const gdb_graph_view &g) {
                return (((unsigned)(g.n + 1) * 4u + 15u) & ~15u) + (((unsigned)g.nnz * 4u + 15u) & ~15u) +
                       (((unsigned)g.nnz * (unsigned)sizeof(edge_t) + 15u) & ~15u);
            };
            auto stage = [&](gdb_graph_view &g, unsigned char *dst) {
                const unsigned b0 = ((unsigned)(g.n + 1) * 4u + 15u) & ~15u, b1 = ((unsigned)g.nnz * 4u + 15u) & ~15u,
                               b2 = ((unsigned)g.nnz * (unsigned)sizeof(edge_t) + 15u) & ~15u;
                gdb_copy16(dst, g.rowptr, b0);
                gdb_copy16(dst + b0, g.rowadj, b1);
                gdb_copy16(dst + b0 + b1, g.edge, b2);
                g.rowptr = reinterpret_cast<const unsigned *>(dst);
                g.rowadj = reinterpret_cast<const unsigned *>(dst + b0);
                g.edge = reinterpret_cast<const edge_t *>(dst + b0 + b1);
            };
            const unsigned need = idx_bytes(g1) + (same ? 0u : idx_bytes(g2));
            if (g1.index16 && g2.index16 && need <= F.smem_bytes) {
                stage(g1, gdb_smem);
                if (same) {
                    g2.rowptr = g1.rowptr;
                    g2.rowadj = g1.rowadj;
                    g2.edge = g1.edge;
                } else {
                    stage(g2, gdb_smem + idx_bytes(g1));
                }
                used = need;
                gdb_group_sync();
            }
        }
        const unsigned Npad = (N + 3u) & ~3u;
        float *vec;
        if ((unsigned long long)used + (unsigned long long)NVEC * Npad * 4ull <= F.smem_bytes) {
            vec = reinterpret_cast<float *>(gdb_smem + used);
        } else {
            vec = F.scratch + (unsigned long long)blockIdx.x * F.scratch_stride;
        }
        float *x = vec, *r = vec + Npad, *p = vec + 2 * Npad, *Ap = vec + 3 * Npad, *diag = vec + 4 * Npad;

        const int n1 = g1.n, n2 = g2.n;
        (void)n1;
        const float Q = 1.0f / (1.0f - F.q);
        const float Q2 = Q * Q;

        // ---- setup: cached diagonal Dx / Vx and right-hand side b = Dx -------
        for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            const float dx = g1.degree[i1] * g2.degree[i2] * Q2;
            const float v = P.node_kernel(g1.node[i1], g2.node[i2]);
            diag[i] = __fdividef(dx, v);
            r[i] = dx;
        }
        int iters = gdb_pcg(P, g1, g2, diag, x, r, p, Ap, F.ftol, s_red, flip);

#if GDB_GRADIENT
        float *y = vec + 5 * Npad;
#endif
#if GDB_GRADIENT && GDB_NODAL == 0
        // ---- adjoint solve  y = A^-1 (p1 (x) p2)  (A is symmetric) -----------
        for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            r[i] = P.p_start(g1.node[i1]) * P.p_start(g2.node[i2]);
        }
        iters += gdb_pcg(P, g1, g2, diag, y, r, p, Ap, F.ftol, s_red, flip);
        gdb_group_sync();  // x, y visible to every thread for the edge sweep
#endif

        if (threadIdx.x == 0) {
            atomicAdd(F.counters + 1, (unsigned long long)iters);
            atomicAdd(F.counters + 2, (unsigned long long)iters * (unsigned long long)g1.nnz * (unsigned long long)g2.nnz);
            atomicAdd(F.counters + 3, (unsigned long long)iters * (unsigned long long)N);
        }

        const unsigned I1 = F.starts[ja] - F.row0, I2 = F.starts[jb] - F.col0;
        const unsigned long long plane = (unsigned long long)F.nX * F.nY;

        // ---- epilogue: apply starting probabilities, write the Gram entry ----
        float norm_rs = 1.f, norm_k = 0.f;
        (void)norm_rs;
        (void)norm_k;
#if GDB_NODAL == 2
        for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            // a self pair is symmetric in exact arithmetic; make it bit-exact
            float xi = 0.5f * (x[i] + x[i2 * n2 + i1]);
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i2]);
#endif
            F.gram[I1 + i1 + i2 * n1] = xi * P.p_start(g1.node[i1]) * P.p_start(g2.node[i2]);
        }
#elif GDB_NODAL == 1 && GDB_DIAGONAL
        for (int i1 = threadIdx.x; i1 < n1; i1 += GDB_BLOCK) {
            float xi = x[i1 * n2 + i1];
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i1]);
#endif
            const float ps = P.p_start(g1.node[i1]);
            F.gram[I1 + i1] = xi * ps * ps;
        }
#elif GDB_NODAL == 1
        for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            float xi = x[i];
#if GDB_SYMMETRIC
            if (ja == jb) xi = 0.5f * (xi + x[i2 * n2 + i1]);  // bit-exact symmetry of self pairs
#endif
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i2]);
#endif
            const float val = xi * P.p_start(g1.node[i1]) * P.p_start(g2.node[i2]);
            F.gram[(unsigned long long)(I1 + i1) + (unsigned long long)(I2 + i2) * F.nX] = val;
#if GDB_SYMMETRIC
            if (ja != jb) F.gram[(unsigned long long)(I2 + i2) + (unsigned long long)(I1 + i1) * F.nX] = val;
#endif
        }
#else
        {
            // per-thread partial sums in double: a thread adds N / GDB_BLOCK terms of one sign and
            // similar size, and a float accumulator rounds them the same way for long stretches
            // (label-free kernels, N = 1.65e6: K off by 6e-5 relative with float partial sums)
            double dsum = 0.0;
            for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
                const int i1 = i / n2, i2 = i - i1 * n2;
                float xi = x[i];
#if GDB_LMIN == 1
                xi -= P.node_kernel(g1.node[i1], g2.node[i2]);
#endif
                dsum += (double)(xi * (P.p_start(g1.node[i1]) * P.p_start(g2.node[i2])));
            }
            float sum = gdb_group_sum((float)dsum, s_red, flip);
#if !GDB_DIAGONAL
            norm_rs = gdb_norm_scale(F, ja, jb);
            sum *= norm_rs;
            norm_k = sum;
#endif
            if (threadIdx.x == 0) {
#if GDB_DIAGONAL
                F.gram[I1] = sum;
#else
                F.gram[(unsigned long long)I1 + (unsigned long long)I2 * F.nX] = sum;
#if GDB_SYMMETRIC
                if (ja != jb) F.gram[(unsigned long long)I2 + (unsigned long long)I1 * F.nX] = sum;
#endif
#endif
            }
        }
#endif

#if GDB_GRADIENT && GDB_NODAL == 0
        // ---- adjoint Jacobian, order [p..., q, node..., edge...] --------------
        // dK/dp_m  = sum (dp1 p2 + p1 dp2) xs          (xs = x - Vx if lmin)
        // dK/dq    = sum y (2Q Dx) - y (2Q Dx / Vx) x
        // dK/dtv_m = sum y x Dx / Vx^2 dVx  [- p1 p2 dVx if lmin]
        // dK/dte_m = sum_{i,j} y_i x_j w1 w2 dEx_ij
        double jac[GDB_NJ];  // per-thread partial sums in double, see the Gram sum above
#pragma unroll
        for (int m = 0; m < GDB_NJ; ++m) jac[m] = 0.0;
        for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            const node_t &u1 = g1.node[i1];
            const node_t &u2 = g2.node[i2];
            const float p1 = P.p_start(u1), p2 = P.p_start(u2);
            const float v = P.node_kernel(u1, u2);
            const float dx = g1.degree[i1] * g2.degree[i2] * Q2;
            const float xi = x[i], yi = y[i];
            float xs = xi;
#if GDB_LMIN == 1
            xs -= v;
#endif
#if GDB_NP > 0
            {
                float d1[GDB_NP], d2[GDB_NP];
                P.p_start.jacobian(u1, d1);
                P.p_start.jacobian(u2, d2);
#pragma unroll
                for (int m = 0; m < GDB_NP; ++m) jac[m] += (double)(fmaf(d1[m], p2, p1 * d2[m]) * xs);
            }
#endif
            jac[GDB_NP] += (double)(2.f * Q * dx * yi * (1.f - __fdividef(xi, v)));
#if GDB_NV > 0
            {
                float dv[GDB_NV];
                P.node_kernel.jacobian(u1, u2, dv);
                const float c = yi * xi * __fdividef(dx, v * v);
#pragma unroll
                for (int m = 0; m < GDB_NV; ++m) {
                    float t = c * dv[m];
#if GDB_LMIN == 1
                    t -= p1 * p2 * dv[m];
#endif
                    jac[GDB_NP + 1 + m] += (double)t;
                }
            }
#endif
#if GDB_NE > 0
            {
                const int r1 = (i1 & 7) * 8, r2 = (i2 & 7) * 8;
                float acc[GDB_NE];
#pragma unroll
                for (int m = 0; m < GDB_NE; ++m) acc[m] = 0.f;
                const unsigned o1_end = g1.trow[(i1 >> 3) + 1], o2_beg = g2.trow[i2 >> 3],
                               o2_end = g2.trow[(i2 >> 3) + 1];
                for (unsigned o1 = g1.trow[i1 >> 3]; o1 < o1_end; ++o1) {
                    const gdb_octile t1 = g1.oct[o1];
                    unsigned m1 = (unsigned)(t1.mask >> r1) & 0xffu;
                    if (!m1) continue;
                    const edge_t *e1 = g1.edge + t1.start + __popcll(t1.mask & ((1ull << r1) - 1ull));
                    const float *x1 = x + (int)t1.tcol * 8 * n2;
                    for (; m1; m1 &= m1 - 1, ++e1) {
                        const float *xrow = x1 + (__ffs(m1) - 1) * n2;
                        for (unsigned o2 = o2_beg; o2 < o2_end; ++o2) {
                            const gdb_octile t2 = g2.oct[o2];
                            unsigned m2 = (unsigned)(t2.mask >> r2) & 0xffu;
                            if (!m2) continue;
                            const edge_t *e2 = g2.edge + t2.start + __popcll(t2.mask & ((1ull << r2) - 1ull));
                            const float *xp = xrow + (int)t2.tcol * 8;
                            for (; m2; m2 &= m2 - 1, ++e2) {
                                float de[GDB_NE];
                                P.edge_kernel.jacobian(e1->label, e2->label, de);
                                float xj = xp[__ffs(m2) - 1];
#if GDB_WEIGHTED
                                xj *= e1->weight * e2->weight;
#endif
#pragma unroll
                                for (int m = 0; m < GDB_NE; ++m) acc[m] = fmaf(de[m], xj, acc[m]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < GDB_NE; ++m) jac[GDB_NP + 1 + GDB_NV + m] += (double)(yi * acc[m]);
            }
#endif
        }
#pragma unroll
        for (int m = 0; m < GDB_NJ; ++m) {
            float s = gdb_group_sum((float)jac[m], s_red, flip);
#if !GDB_DIAGONAL
            s = gdb_norm_grad(F, ja, jb, m, norm_rs, norm_k, s);
#endif
            if (threadIdx.x == 0) {
#if GDB_DIAGONAL
                F.grad[(unsigned long long)I1 + (unsigned long long)m * F.nX] = s;
#else
                F.grad[(unsigned long long)I1 + (unsigned long long)I2 * F.nX + m * plane] = s;
#if GDB_SYMMETRIC
                if (ja != jb) F.grad[(unsigned long long)I2 + (unsigned long long)I1 * F.nX + m * plane] = s;
#endif
#endif
            }
        }
#endif
#if GDB_GRADIENT && GDB_NODAL != 0
        // ---- nodal Jacobian by forward sensitivities ------------------------------
        // R_i = xs_i p1 p2.  d/dp_m is explicit; for q, node and edge parameters
        //     dx/dt = A^-1 (db/dt - dA/dt x)
        // is one more solve with the same operator per parameter (the reference
        // instead re-solves twice per parameter for a central difference,
        // template.cu:226-418):
        //   q:      rhs = 2Q Dx (1 - x / Vx)
        //   node m: rhs = Dx / Vx^2 dVx_m x          (+ lmin: R -= dVx_m p1 p2)
        //   edge m: rhs = (dW_m) x = sum_j w1 w2 dkE_m x_j
        {
            auto write_nodal = [&](int i, int i1, int i2, int m, float val) {
#if GDB_NODAL == 2
                if (true) F.grad[(unsigned long long)(I1 + i1 + i2 * n1) + (unsigned long long)m * F.nX] = val;
#elif GDB_DIAGONAL
                if (i1 == i2) F.grad[(unsigned long long)(I1 + i1) + (unsigned long long)m * F.nX] = val;
#else
                F.grad[(unsigned long long)(I1 + i1) + (unsigned long long)(I2 + i2) * F.nX + m * plane] = val;
#if GDB_SYMMETRIC
                if (ja != jb) F.grad[(unsigned long long)(I2 + i2) + (unsigned long long)(I1 + i1) * F.nX + m * plane] = val;
#endif
#endif
                (void)i;
            };
#if GDB_NP > 0
            for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
                const int i1 = i / n2, i2 = i - i1 * n2;
                const node_t &u1 = g1.node[i1];
                const node_t &u2 = g2.node[i2];
                float xs = x[i];
#if GDB_NODAL == 2 || GDB_SYMMETRIC
                if (ja == jb) xs = 0.5f * (xs + x[i2 * n2 + i1]);  // self pair: bit-exact symmetry
#endif
#if GDB_LMIN == 1
                xs -= P.node_kernel(u1, u2);
#endif
                float d1[GDB_NP], d2[GDB_NP];
                P.p_start.jacobian(u1, d1);
                P.p_start.jacobian(u2, d2);
                const float p1 = P.p_start(u1), p2 = P.p_start(u2);
#pragma unroll
                for (int m = 0; m < GDB_NP; ++m) write_nodal(i, i1, i2, m, xs * __fadd_rn(__fmul_rn(d1[m], p2), __fmul_rn(p1, d2[m])));  // no FMA: symmetric under swap
            }
#endif
            int extra_iters = 0;
#pragma unroll 1
            for (int m = GDB_NP; m < GDB_NJ; ++m) {
                gdb_group_sync();  // x stable, r free
                for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
                    const int i1 = i / n2, i2 = i - i1 * n2;
                    const node_t &u1 = g1.node[i1];
                    const node_t &u2 = g2.node[i2];
                    const float dx = g1.degree[i1] * g2.degree[i2] * Q2;
                    float rhs = 0.f;
                    if (m == GDB_NP) {
                        rhs = 2.f * Q * dx * (1.f - __fdividef(x[i], P.node_kernel(u1, u2)));
                    }
#if GDB_NV > 0
                    else if (m < GDB_NP + 1 + GDB_NV) {
                        float dv[GDB_NV];
                        P.node_kernel.jacobian(u1, u2, dv);
                        float dvm = 0.f;
#pragma unroll
                        for (int k = 0; k < GDB_NV; ++k) dvm = (k == m - GDB_NP - 1) ? dv[k] : dvm;
                        const float v = P.node_kernel(u1, u2);
                        rhs = __fdividef(dx, v * v) * dvm * x[i];
                    }
#endif
#if GDB_NE > 0
                    else {
                        const int r1 = (i1 & 7) * 8, r2 = (i2 & 7) * 8;
                        const unsigned o1_end = g1.trow[(i1 >> 3) + 1], o2_beg = g2.trow[i2 >> 3],
                                       o2_end = g2.trow[(i2 >> 3) + 1];
                        for (unsigned o1 = g1.trow[i1 >> 3]; o1 < o1_end; ++o1) {
                            const gdb_octile t1 = g1.oct[o1];
                            unsigned m1 = (unsigned)(t1.mask >> r1) & 0xffu;
                            if (!m1) continue;
                            const edge_t *e1 = g1.edge + t1.start + __popcll(t1.mask & ((1ull << r1) - 1ull));
                            const float *x1 = x + (int)t1.tcol * 8 * n2;
                            for (; m1; m1 &= m1 - 1, ++e1) {
                                const float *xrow = x1 + (__ffs(m1) - 1) * n2;
                                for (unsigned o2 = o2_beg; o2 < o2_end; ++o2) {
                                    const gdb_octile t2 = g2.oct[o2];
                                    unsigned m2 = (unsigned)(t2.mask >> r2) & 0xffu;
                                    if (!m2) continue;
                                    const edge_t *e2 = g2.edge + t2.start + __popcll(t2.mask & ((1ull << r2) - 1ull));
                                    const float *xp = xrow + (int)t2.tcol * 8;
                                    for (; m2; m2 &= m2 - 1, ++e2) {
                                        float de[GDB_NE];
                                        P.edge_kernel.jacobian(e1->label, e2->label, de);
                                        float dem = 0.f;
#pragma unroll
                                        for (int k = 0; k < GDB_NE; ++k) dem = (k == m - GDB_NP - 1 - GDB_NV) ? de[k] : dem;
#if GDB_WEIGHTED
                                        dem *= e1->weight * e2->weight;
#endif
                                        rhs = fmaf(dem, xp[__ffs(m2) - 1], rhs);
                                    }
                                }
                            }
                        }
                    }
#endif
                    r[i] = rhs;
                }
                extra_iters += gdb_pcg(P, g1, g2, diag, y, r, p, Ap, F.ftol, s_red, flip);
                gdb_group_sync();
                for (int i = threadIdx.x; i < (int)N; i += GDB_BLOCK) {
                    const int i1 = i / n2, i2 = i - i1 * n2;
                    const node_t &u1 = g1.node[i1];
                    const node_t &u2 = g2.node[i2];
                    float val = y[i];
#if GDB_NODAL == 2 || GDB_SYMMETRIC
                    if (ja == jb) val = 0.5f * (val + y[i2 * n2 + i1]);  // self pair: bit-exact symmetry
#endif
#if GDB_LMIN == 1 && GDB_NV > 0
                    if (m > GDB_NP && m < GDB_NP + 1 + GDB_NV) {
                        float dv[GDB_NV];
                        P.node_kernel.jacobian(u1, u2, dv);
#pragma unroll
                        for (int k = 0; k < GDB_NV; ++k) val -= (k == m - GDB_NP - 1) ? dv[k] : 0.f;
                    }
#endif
                    write_nodal(i, i1, i2, m, val * P.p_start(u1) * P.p_start(u2));
                }
            }
            if (threadIdx.x == 0) {
                atomicAdd(F.counters + 1, (unsigned long long)extra_iters);
                atomicAdd(F.counters + 2, (unsigned long long)extra_iters * (unsigned long long)g1.nnz * (unsigned long long)g2.nnz);
                atomicAdd(F.counters + 3, (unsigned long long)extra_iters * (unsigned long long)N);
            }
        }
#endif
        (void)plane;
    }
}
#endif  // GDB_BUILD_MASK & 1
