// mlgk_small.cuh -- shared-memory-resident solver for small graph pairs
// (everything of a pair fits in one CTA's shared memory: both graph blobs, the
// cached edge-kernel products, and the CG vectors).  This is the kernel behind
// the BASELINE configurations C1, C2, C3 and C5 (molecular graphs of ~20
// nodes).  Same algorithm and same results contract as mlgk_solve
// (mlgk_solver.cuh); it replaces reference
// graphdot/cpp/marginalized_kernel.h:189-490 (compute), :492-804
// (compute_duo) and :806-997 (derivative) for pairs in this regime.
//
// What is different from the general kernel, and why:
//  * W = w1 w2 kE(e1, e2) is evaluated ONCE per pair for all nnz1 x nnz2
//    element pairs (perfectly balanced, no divergence) and kept in shared
//    memory; every CG iteration of both solves then costs one shared load per
//    product instead of re-evaluating the edge microkernel (the reference
//    re-evaluates it for every product in every iteration,
//    marginalized_kernel.h:299-300, :346).
//  * the matvec is organised by "workers" = (tile row T1 of G1) x (column i2
//    of G2).  A worker walks the compact elements of the octiles in tile row
//    T1 (one contiguous range, uniform across the lanes that share T1) and,
//    per element, the neighbours of i2 from the row index of G2; it owns the
//    outputs (rows of T1, column i2), so there are no atomics and no races.
//  * with gradients, the value system (rhs Dx) and the adjoint system (rhs
//    p1 (x) p2) are solved TOGETHER on float2 vectors: one W load and one
//    64-bit vector load feed two FMAs.  Each system keeps its own CG scalars
//    and convergence flag (the reference shares alpha/beta between the two
//    stacked systems, marginalized_kernel.h:721-772).
#pragma once

#if GDB_GRADIENT
typedef float2 gv_t;
#define GV_N 2
__device__ __forceinline__ gv_t gv_make(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ gv_t gv_fma(float a, gv_t b, gv_t c) { return make_float2(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y)); }
__device__ __forceinline__ gv_t gv_fma2(gv_t a, gv_t b, gv_t c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ gv_t gv_sub(gv_t a, gv_t b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ gv_t gv_scale(float a, gv_t b) { return make_float2(a * b.x, a * b.y); }
__device__ __forceinline__ gv_t gv_neg(gv_t a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float gv_get(gv_t a, int k) { return k ? a.y : a.x; }
#else
typedef float gv_t;
#define GV_N 1
__device__ __forceinline__ gv_t gv_make(float a, float) { return a; }
__device__ __forceinline__ gv_t gv_fma(float a, gv_t b, gv_t c) { return fmaf(a, b, c); }
__device__ __forceinline__ gv_t gv_fma2(gv_t a, gv_t b, gv_t c) { return fmaf(a, b, c); }
__device__ __forceinline__ gv_t gv_sub(gv_t a, gv_t b) { return a - b; }
__device__ __forceinline__ gv_t gv_scale(float a, gv_t b) { return a * b; }
__device__ __forceinline__ gv_t gv_neg(gv_t a) { return -a; }
__device__ __forceinline__ float gv_get(gv_t a, int) { return a; }
#endif

// Group sum of K values at once (one barrier); fixed summation order.
template<int K> __device__ __forceinline__ void gdb_group_sum_n(float (&v)[K], float *red, int &flip) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = gdb_warp_sum(v[k]);
#if GDB_BLOCK > 32
    float *buf = red + flip * (GDB_WARPS * 4);
    flip ^= 1;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) buf[(threadIdx.x >> 5) * 4 + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < GDB_WARPS; ++w) t += buf[w * 4 + k];
        v[k] = t;
    }
#else
    __syncwarp();  // orders the lanes' shared-memory writes like the barrier above
#endif
}

struct gdb_small_graph {
    const float *degree;
    const node_t *node;
    const edge_t *edge;
    const unsigned *emeta, *rowptr, *rowadj, *tileelem;
    int n, nnz, n_tile;
};

__device__ __forceinline__ gdb_small_graph gdb_small_view(const unsigned char *base) {
    const gdb_graph_hdr *h = reinterpret_cast<const gdb_graph_hdr *>(base);
    gdb_small_graph v;
    v.degree = reinterpret_cast<const float *>(base + h->off_degree);
    v.node = reinterpret_cast<const node_t *>(base + h->off_node);
    v.edge = reinterpret_cast<const edge_t *>(base + h->off_edge);
    v.emeta = reinterpret_cast<const unsigned *>(base + h->off_emeta);
    v.rowptr = reinterpret_cast<const unsigned *>(base + h->off_rowptr);
    v.rowadj = reinterpret_cast<const unsigned *>(base + h->off_rowadj);
    v.tileelem = reinterpret_cast<const unsigned *>(base + h->off_tileelem);
    v.n = h->n_node;
    v.nnz = h->nnz;
    v.n_tile = h->n_tile;
    return v;
}

extern "C" __global__ void __launch_bounds__(GDB_BLOCK, GDB_MIN_BLOCKS)
    mlgk_solve_small(const __grid_constant__ gdb_params P) {
    extern __shared__ __align__(16) unsigned char gdb_smem[];
    __shared__ unsigned long long s_job;
    __shared__ float s_red[2 * 4 * (GDB_WARPS > 0 ? GDB_WARPS : 1)];
    int flip = 0;
    const gdb_params_fixed &F = P.f;

    while (true) {
        gdb_group_sync();  // previous job's shared memory is dead
        if (threadIdx.x == 0) s_job = atomicAdd(F.counters, 1ull);
        gdb_group_sync();
        const unsigned long long job = s_job;
        if (job >= F.n_jobs) break;
        unsigned ja, jb;
        gdb_decode_job(F, job, ja, jb);
        const gdb_graph_ref ref1 = F.graphs[ja], ref2 = F.graphs[jb];
        const bool same = (ja == jb);

        // ---- stage both graphs ------------------------------------------------
        gdb_copy16(gdb_smem, ref1.blob, ref1.bytes);
        unsigned used = ref1.bytes;
        const unsigned char *base2 = gdb_smem;
        if (!same) {
            gdb_copy16(gdb_smem + used, ref2.blob, ref2.bytes);
            base2 = gdb_smem + used;
            used += ref2.bytes;
        }
        gdb_group_sync();
        const gdb_small_graph g1 = gdb_small_view(gdb_smem), g2 = gdb_small_view(base2);
        const int n1 = g1.n, n2 = g2.n, N = n1 * n2, nnz1 = g1.nnz, nnz2 = g2.nnz;
        const int Npad = (N + 3) & ~3;
        float *W = reinterpret_cast<float *>(gdb_smem + used);
        float *diag = W + ((nnz1 * nnz2 + 3) & ~3);
        gv_t *x = reinterpret_cast<gv_t *>(diag + Npad);
        gv_t *r = x + Npad, *p = r + Npad, *Ap = p + Npad;

        // ---- W = w1 w2 kE(e1, e2), once per pair ---------------------------------
        for (int idx = threadIdx.x; idx < nnz1 * nnz2; idx += GDB_BLOCK) {
            const int e1 = idx / nnz2, e2 = idx - e1 * nnz2;
            W[idx] = gdb_edge_value(P, g1.edge[e1], g2.edge[e2]);
        }
        // ---- diagonal, right-hand sides, CG start ---------------------------------
        const float Q = 1.0f / (1.0f - F.q), Q2 = Q * Q;
        float rho[GV_N];
#pragma unroll
        for (int k = 0; k < GV_N; ++k) rho[k] = 0.f;
        for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            const node_t &u1 = g1.node[i1];
            const node_t &u2 = g2.node[i2];
            const float dx = g1.degree[i1] * g2.degree[i2] * Q2;
            const float v = P.node_kernel(u1, u2);
            const float d = __fdividef(dx, v);
            diag[i] = d;
#if GDB_GRADIENT
            const gv_t ri = gv_make(dx, P.p_start(u1) * P.p_start(u2));
#else
            const gv_t ri = gv_make(dx, 0.f);
#endif
            const gv_t z = gv_scale(__fdividef(1.0f, d), ri);
            x[i] = gv_make(0.f, 0.f);
            r[i] = ri;
            p[i] = z;
#pragma unroll
            for (int k = 0; k < GV_N; ++k) rho[k] = fmaf(gv_get(ri, k), gv_get(z, k), rho[k]);
        }
        gdb_group_sum_n(rho, s_red, flip);
        gdb_group_sync();  // W, p complete

        // ---- Jacobi-PCG, both systems at once ---------------------------------------
        const float thresh = F.ftol * (float)N;
        bool active[GV_N];
#pragma unroll
        for (int k = 0; k < GV_N; ++k) active[k] = rho[k] != 0.f;
        int iters = 0;  // summed over the systems that were still active
        const int n_worker = g1.n_tile * n2;
        for (int it = 0; it < N; ++it) {
            bool any = false;
#pragma unroll
            for (int k = 0; k < GV_N; ++k) any |= active[k];
            if (!any) break;
#pragma unroll
            for (int k = 0; k < GV_N; ++k) iters += active[k] ? 1 : 0;
            // matvec: Ap = diag p - W p
            float pAp[GV_N];
#pragma unroll
            for (int k = 0; k < GV_N; ++k) pAp[k] = 0.f;
            for (int w = threadIdx.x; w < n_worker; w += GDB_BLOCK) {
                const int T1 = w / n2, i2 = w - T1 * n2;
                const int row_end = min(8 * T1 + 8, n1);
                for (int i1 = 8 * T1; i1 < row_end; ++i1) Ap[i1 * n2 + i2] = gv_scale(diag[i1 * n2 + i2], p[i1 * n2 + i2]);
                const unsigned kbeg = g2.rowptr[i2], kend = g2.rowptr[i2 + 1];
                const unsigned eend = g1.tileelem[T1 + 1];
                for (unsigned e1 = g1.tileelem[T1]; e1 < eend; ++e1) {
                    const unsigned m = g1.emeta[e1];
                    const float *Wrow = W + e1 * nnz2;
                    const gv_t *prow = p + (m >> 16) * n2;
                    gv_t acc = gv_make(0.f, 0.f);
                    for (unsigned k = kbeg; k < kend; ++k) {
                        const unsigned a = g2.rowadj[k];
                        acc = gv_fma(Wrow[a >> 16], prow[a & 0xffffu], acc);
                    }
                    gv_t *dst = Ap + (m & 0xffffu) * n2 + i2;
                    *dst = gv_sub(*dst, acc);
                }
                for (int i1 = 8 * T1; i1 < row_end; ++i1) {
                    const gv_t pv = p[i1 * n2 + i2], av = Ap[i1 * n2 + i2];
#pragma unroll
                    for (int k = 0; k < GV_N; ++k) pAp[k] = fmaf(gv_get(pv, k), gv_get(av, k), pAp[k]);
                }
            }
            gdb_group_sum_n(pAp, s_red, flip);  // barrier: Ap complete
            float alpha[GV_N];
#pragma unroll
            for (int k = 0; k < GV_N; ++k) {
                if (pAp[k] == 0.f) active[k] = false;
                alpha[k] = active[k] ? __fdividef(rho[k], pAp[k]) : 0.f;
            }
            const gv_t al = gv_make(alpha[0], alpha[GV_N - 1]);
            float s[2 * GV_N];
#pragma unroll
            for (int k = 0; k < 2 * GV_N; ++k) s[k] = 0.f;
            for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
                x[i] = gv_fma2(al, p[i], x[i]);
                const gv_t ri = gv_fma2(gv_neg(al), Ap[i], r[i]);
                r[i] = ri;
                const float dinv = __fdividef(1.0f, diag[i]);
#pragma unroll
                for (int k = 0; k < GV_N; ++k) {
                    const float rk = gv_get(ri, k);
                    s[2 * k] = fmaf(rk, rk, s[2 * k]);
                    s[2 * k + 1] = fmaf(rk * dinv, rk, s[2 * k + 1]);
                }
            }
            gdb_group_sum_n(s, s_red, flip);
            float beta[GV_N];
#pragma unroll
            for (int k = 0; k < GV_N; ++k) {
                if (active[k] && sqrtf(s[2 * k]) < thresh) active[k] = false;
                beta[k] = active[k] ? __fdividef(s[2 * k + 1], rho[k]) : 0.f;
                if (active[k]) rho[k] = s[2 * k + 1];
                if (rho[k] == 0.f) active[k] = false;
            }
            const gv_t be = gv_make(beta[0], beta[GV_N - 1]);
            for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
                const gv_t z = gv_scale(__fdividef(1.0f, diag[i]), r[i]);
                p[i] = gv_fma2(be, p[i], z);
            }
            gdb_group_sync();  // p complete before the next matvec
        }

        if (threadIdx.x == 0) {
            atomicAdd(F.counters + 1, (unsigned long long)iters);
            atomicAdd(F.counters + 2, (unsigned long long)iters * (unsigned long long)nnz1 * (unsigned long long)nnz2);
            atomicAdd(F.counters + 3, (unsigned long long)iters * (unsigned long long)N);
        }

        const unsigned I1 = F.starts[ja] - F.row0, I2 = F.starts[jb] - F.col0;
        const unsigned long long plane = (unsigned long long)F.nX * F.nY;
        (void)plane;

        // ---- epilogue: starting probabilities, Gram entry ----------------------------
#if GDB_NODAL == 2
        for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            float xi = 0.5f * (gv_get(x[i], 0) + gv_get(x[i2 * n2 + i1], 0));  // self pair: bit-exact symmetry
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i2]);
#endif
            F.gram[I1 + i1 + i2 * n1] = xi * P.p_start(g1.node[i1]) * P.p_start(g2.node[i2]);
        }
#elif GDB_NODAL == 1 && GDB_DIAGONAL
        for (int i1 = threadIdx.x; i1 < n1; i1 += GDB_BLOCK) {
            float xi = gv_get(x[i1 * n2 + i1], 0);
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i1]);
#endif
            const float ps = P.p_start(g1.node[i1]);
            F.gram[I1 + i1] = xi * ps * ps;
        }
#elif GDB_NODAL == 1
        for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            float xi = gv_get(x[i], 0);
#if GDB_SYMMETRIC
            if (same) xi = 0.5f * (xi + gv_get(x[i2 * n2 + i1], 0));  // bit-exact symmetry of self pairs
#endif
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i2]);
#endif
            const float val = xi * P.p_start(g1.node[i1]) * P.p_start(g2.node[i2]);
            F.gram[(unsigned long long)(I1 + i1) + (unsigned long long)(I2 + i2) * F.nX] = val;
#if GDB_SYMMETRIC
            if (!same) F.gram[(unsigned long long)(I2 + i2) + (unsigned long long)(I1 + i1) * F.nX] = val;
#endif
        }
#else
        // graph level: K = sum xs p1 p2; with gradients also the node-side Jacobian terms
        {
            float acc[1 + (GDB_GRADIENT ? GDB_NP + 1 + GDB_NV : 0)];
#pragma unroll
            for (int m = 0; m < (int)(sizeof(acc) / sizeof(float)); ++m) acc[m] = 0.f;
            for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
                const int i1 = i / n2, i2 = i - i1 * n2;
                const node_t &u1 = g1.node[i1];
                const node_t &u2 = g2.node[i2];
                const float p1 = P.p_start(u1), p2 = P.p_start(u2);
                const float xi = gv_get(x[i], 0);
                float xs = xi;
#if GDB_LMIN == 1 || GDB_GRADIENT
                const float v = P.node_kernel(u1, u2);
#endif
#if GDB_LMIN == 1
                xs -= v;
#endif
                acc[0] = fmaf(xs, p1 * p2, acc[0]);
#if GDB_GRADIENT
                // dK/dp_m  = sum (dp1 p2 + p1 dp2) xs
                // dK/dq    = sum y (2Q Dx) (1 - x / Vx)
                // dK/dtv_m = sum y x Dx / Vx^2 dVx  [- p1 p2 dVx if lmin]
                const float yi = gv_get(x[i], 1);
                const float dx = diag[i] * v;
#if GDB_NP > 0
                {
                    float d1[GDB_NP], d2[GDB_NP];
                    P.p_start.jacobian(u1, d1);
                    P.p_start.jacobian(u2, d2);
#pragma unroll
                    for (int m = 0; m < GDB_NP; ++m) acc[1 + m] = fmaf(fmaf(d1[m], p2, p1 * d2[m]), xs, acc[1 + m]);
                }
#endif
                acc[1 + GDB_NP] += 2.f * Q * dx * yi * (1.f - __fdividef(xi, v));
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
                        acc[2 + GDB_NP + m] += t;
                    }
                }
#endif
#endif
            }
#if GDB_GRADIENT && GDB_NE > 0
            // dK/dte_m = sum_{i,j} y_i x_j w1 w2 dkE_m(e1, e2): one balanced pass
            // over all element pairs
            float eacc[GDB_NE];
#pragma unroll
            for (int m = 0; m < GDB_NE; ++m) eacc[m] = 0.f;
            for (int idx = threadIdx.x; idx < nnz1 * nnz2; idx += GDB_BLOCK) {
                const int e1 = idx / nnz2, e2 = idx - e1 * nnz2;
                const unsigned m1 = g1.emeta[e1], m2 = g2.emeta[e2];
                const float yi = gv_get(x[(m1 & 0xffffu) * n2 + (m2 & 0xffffu)], 1);
                float xj = gv_get(x[(m1 >> 16) * n2 + (m2 >> 16)], 0);
                const edge_t &a = g1.edge[e1];
                const edge_t &b = g2.edge[e2];
#if GDB_WEIGHTED
                xj *= a.weight * b.weight;
#endif
                float de[GDB_NE];
                P.edge_kernel.jacobian(a.label, b.label, de);
#pragma unroll
                for (int m = 0; m < GDB_NE; ++m) eacc[m] = fmaf(de[m] * yi, xj, eacc[m]);
            }
#endif
            constexpr int NACC = (int)(sizeof(acc) / sizeof(float));
#pragma unroll
            for (int m0 = 0; m0 < NACC; m0 += 4) {
                float part[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) part[k] = (m0 + k < NACC) ? acc[m0 + k] : 0.f;
                gdb_group_sum_n(part, s_red, flip);
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
                gdb_group_sum_n(part, s_red, flip);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (m0 + k < GDB_NE) eacc[m0 + k] = part[k];
            }
#endif
            if (threadIdx.x == 0) {
#if GDB_DIAGONAL
                F.gram[I1] = acc[0];
#else
                F.gram[(unsigned long long)I1 + (unsigned long long)I2 * F.nX] = acc[0];
#if GDB_SYMMETRIC
                if (!same) F.gram[(unsigned long long)I2 + (unsigned long long)I1 * F.nX] = acc[0];
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
                    F.grad[(unsigned long long)I1 + (unsigned long long)I2 * F.nX + m * plane] = val;
#if GDB_SYMMETRIC
                    if (!same) F.grad[(unsigned long long)I2 + (unsigned long long)I1 * F.nX + m * plane] = val;
#endif
#endif
                }
#endif
            }
        }
#endif
    }
}
