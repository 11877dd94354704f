// mlgk_small.cuh -- shared-memory / register-resident solver for small graph
// pairs: both graph blobs, the cached edge-kernel products W and one vector fit
// in a CTA's shared memory and the CG state fits in registers.  This is the
// kernel behind the BASELINE configurations C1, C2, C3 and C5 (molecular
// graphs of ~20 nodes).  Same algorithm and results contract as mlgk_solve
// (mlgk_solver.cuh); it replaces reference
// graphdot/cpp/marginalized_kernel.h:189-490 (compute), :492-804
// (compute_duo) and :806-997 (derivative) for pairs in this regime.
//
// Design (measurements and the variants that lost: DESIGN.md sections 4.1, 10):
//  * W = w1 w2 kE(e1, e2) is evaluated ONCE per pair for all nnz1 x nnz2
//    element pairs (balanced, divergence-free) into shared memory, zero where a
//    lane has fewer neighbours than slots, so that the matvec needs no
//    predicates.  The reference re-evaluates the edge microkernel for every
//    product in every CG iteration (marginalized_kernel.h:299-300, :346).
//  * workers = (block of <= GDB_RPW rows of G1, dealt evenly to ALL warps) x
//    (column of G2 = lane), one or a few per thread.  A worker owns the
//    product-graph elements (its rows, its column): their x, r, A p and diagonal
//    live in REGISTERS for the whole solve; only the search direction p, which
//    neighbours gather, is in shared memory.
//  * virtual lanes: lane position p < n2 owns the column lane_map[p] of G2 (the
//    packer's degree-sorted order, gdb_pack.cpp) and its first GDB_ADJ neighbour
//    slots, kept in registers as byte offsets; a column with more neighbours gets
//    HELPER lanes (n2, n2 + 1, ...: the lanes a warp would otherwise idle), one
//    per further chunk of GDB_ADJ slots, and receives their partial sums by one
//    shuffle per row.  What finds no lane (rare) is a compact overflow region
//    walked by the owner.  One matvec step is
//        LDS.64 step table (W row, p row), LDS.64/128 W, GDB_ADJ x LDS p, FMAs.
//    No atomics, fixed summation order => bit-reproducible.
//  * K and its Jacobian are symmetric in the two graphs, so per pair the graph
//    whose virtual columns fit the lanes (then the larger one) provides the
//    columns and the other one the rows.
//  * with gradients the value system (rhs Dx) and the adjoint system (rhs
//    p1 (x) p2) are solved together on float2 data: one W load and one 64-bit
//    load feed two FMAs.  Each system has its own CG scalars and convergence
//    flag (the reference shares alpha/beta between the stacked systems,
//    marginalized_kernel.h:721-772).
//  * measured and rejected in round 2: a software-
//    pipelined element loop (+0.4 %, within noise) and the symmetrically scaled
//    unit-diagonal system (-2.6 %: the scaling of W and the extra shared-memory
//    vector cost more than the divisions they save, and diag^-1/2 does not exist
//    for the negative degrees that the reference's tests allow);
//  * three barriers per CG iteration (two reductions + publish p); the K dot
//    products of a reduction share one shuffle butterfly and are joined through
//    a double-buffered shared-memory slot.
//  * the next pair's blobs are prefetched by TMA bulk copies (cp.async.bulk +
//    mbarrier, double buffered) while the current pair is solved.
//  * the kernel is bound by instruction issue and shared-memory latency, and its
//    CG loop must stay inside the instruction cache: heavy one-time code (node
//    kernel, Jacobians) runs in rolled loops that talk to the row registers
//    through shared memory, rare paths are out of line, and addresses that the
//    compiler would re-derive at every access are pinned in registers.
//
// Macros from the generated header: GDB_BLOCK, GDB_WPT (workers per thread:
// 32 * GDB_WPT >= nodes), GDB_RPW (rows per warp: GDB_RPW * warps >= nodes),
// GDB_ADJ (neighbour slots per lane, 2 or 4), GDB_MIN_BLOCKS_SMALL.
#pragma once

#ifndef GDB_RPW
#define GDB_RPW 8  // rows of G1 per warp (register arrays are sized by it)
#endif
#ifndef GDB_WPT
#define GDB_WPT 1
#endif
#ifndef GDB_ADJ
#define GDB_ADJ 4  // neighbour slots per (virtual) lane: 2 or 4
#endif
#ifndef GDB_K1_UNROLL
#define GDB_K1_UNROLL 1  // unroll factor of the element loop of a row (A-B hook)
#endif
#define GDB_PRAGMA_(x) _Pragma(#x)
#define GDB_UNROLL(n) GDB_PRAGMA_(unroll n)
#ifndef GDB_ROLL_ROWS
#define GDB_ROLL_ROWS 0  // 1: matvec rolled over the rows, W p handed over through shared memory
#endif
#ifndef GDB_TMA_STAGE
#define GDB_TMA_STAGE 1  // 0: synchronous uint4 staging (tuning / A-B hook)
#endif
#ifdef GDB_SMALL_MINB  // tuning hook: -DGDB_SMALL_MINB=<resident CTAs per SM>
#undef GDB_MIN_BLOCKS_SMALL
#define GDB_MIN_BLOCKS_SMALL GDB_SMALL_MINB
#endif

#if GDB_GRADIENT
typedef float2 gv_t;
#define GV_N 2
__device__ __forceinline__ gv_t gv_make(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ gv_t gv_fma(float a, gv_t b, gv_t c) { return make_float2(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y)); }
__device__ __forceinline__ gv_t gv_fma2(gv_t a, gv_t b, gv_t c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ gv_t gv_scale(float a, gv_t b) { return make_float2(a * b.x, a * b.y); }
__device__ __forceinline__ gv_t gv_neg(gv_t a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ gv_t gv_add(gv_t a, gv_t b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float gv_get(gv_t a, int k) { return k ? a.y : a.x; }
#else
typedef float gv_t;
#define GV_N 1
__device__ __forceinline__ gv_t gv_make(float a, float) { return a; }
__device__ __forceinline__ gv_t gv_fma(float a, gv_t b, gv_t c) { return fmaf(a, b, c); }
__device__ __forceinline__ gv_t gv_fma2(gv_t a, gv_t b, gv_t c) { return fmaf(a, b, c); }
__device__ __forceinline__ gv_t gv_scale(float a, gv_t b) { return a * b; }
__device__ __forceinline__ gv_t gv_neg(gv_t a) { return -a; }
__device__ __forceinline__ gv_t gv_add(gv_t a, gv_t b) { return a + b; }
__device__ __forceinline__ float gv_get(gv_t a, int) { return a; }
#endif

// nodal Jacobians (forward sensitivities) run extra solve rounds
#define GDB_NGRAD (GDB_GRADIENT && GDB_NODAL != 0)
// byte offset added to the W-row addresses of a matvec: a compile-time zero for the CG
// matvec, the distance to the dW buffer for the right-hand sides of edge sensitivities
struct gdb_wzero {
    __device__ __forceinline__ unsigned operator()() const { return 0u; }
};
struct gdb_wdelta {
    unsigned bytes;
    __device__ __forceinline__ unsigned operator()() const { return bytes; }
};

// ---- mbarrier / TMA bulk copy (cp.async.bulk) helpers --------------------------
__device__ __forceinline__ unsigned gdb_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gdb_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gdb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gdb_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void gdb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void gdb_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gdb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gdb_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     gdb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(gdb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void gdb_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "GDB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GDB_DONE_%=;\n\t"
        "bra GDB_WAIT_%=;\n\t"
        "GDB_DONE_%=:\n\t"
        "}" ::"r"(gdb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Group sum of K <= 4 values at once (one barrier); fixed summation order.
// Within a warp the K sums share one butterfly: at the first steps every lane
// passes on half of its values and keeps the other half (6 shuffles for K = 4
// instead of 20); lane 8 k ends up with the warp total of value k (K = 4), lane
// 16 k (K = 2).
template<int K> __device__ __forceinline__ void gdb_group_sum_n(float (&v)[K], float *red, int &flip) {
    static_assert(K == 1 || K == 2 || K == 4, "gdb_group_sum_n: K must be 1, 2 or 4");
    const unsigned lane = threadIdx.x & 31u;
    float keep;
    if constexpr (K == 4) {
        const bool hi = lane & 16u;
        float ka = hi ? v[2] : v[0], kb = hi ? v[3] : v[1];
        ka += __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[2], 16);
        kb += __shfl_xor_sync(0xffffffffu, hi ? v[1] : v[3], 16);
        const bool hi8 = lane & 8u;
        keep = (hi8 ? kb : ka) + __shfl_xor_sync(0xffffffffu, hi8 ? ka : kb, 8);
        keep += __shfl_xor_sync(0xffffffffu, keep, 4);
        keep += __shfl_xor_sync(0xffffffffu, keep, 2);
        keep += __shfl_xor_sync(0xffffffffu, keep, 1);
    } else if constexpr (K == 2) {
        const bool hi = lane & 16u;
        keep = (hi ? v[K - 1] : v[0]) + __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[K - 1], 16);
        keep += __shfl_xor_sync(0xffffffffu, keep, 8);
        keep += __shfl_xor_sync(0xffffffffu, keep, 4);
        keep += __shfl_xor_sync(0xffffffffu, keep, 2);
        keep += __shfl_xor_sync(0xffffffffu, keep, 1);
    } else {
        keep = gdb_warp_sum(v[0]);
    }
#if GDB_BLOCK > 32
    float *buf = red + flip * (GDB_WARPS * 4);
    flip ^= 1;
    if ((lane & (32u / K - 1u)) == 0u) buf[(threadIdx.x >> 5) * 4 + lane / (32u / K)] = keep;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < GDB_WARPS; ++w) t += buf[w * 4 + k];
        v[k] = t;
    }
#else
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = __shfl_sync(0xffffffffu, keep, k * (32 / K));
    __syncwarp();  // orders the lanes' shared-memory accesses like the barrier above
#endif
}

// q = a / d, r = a % d for a < 2^24 through one float multiply (+ fix-up)
__device__ __forceinline__ void gdb_divmod(unsigned a, unsigned d, float inv_d, unsigned &q, unsigned &r) {
    q = (unsigned)(__uint2float_rz(a) * inv_d);
    r = a - q * d;
    if ((int)r < 0) {
        --q;
        r += d;
    } else if (r >= d) {
        ++q;
        r -= d;
    }
}

// Shared-memory loads by 32-bit shared-window address (no generic-address
// arithmetic, always LDS).
__device__ __forceinline__ gv_t gdb_lds_gv(unsigned addr) {
#if GDB_GRADIENT
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
#else
    float x;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
    return x;
#endif
}
__device__ __forceinline__ float2 gdb_lds_f2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 gdb_lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// predicated variants: a lane whose predicate is off issues no request; the load returns zeros
__device__ __forceinline__ gv_t gdb_lds_gv_if(unsigned addr, bool on) {
#if GDB_GRADIENT
    float x = 0.f, y = 0.f;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q ld.shared.v2.f32 {%0, %1}, [%2]; }"
                 : "+f"(x), "+f"(y)
                 : "r"(addr), "r"((unsigned)on));
    return make_float2(x, y);
#else
    float x = 0.f;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.shared.f32 %0, [%1]; }" : "+f"(x) : "r"(addr), "r"((unsigned)on));
    return x;
#endif
}
__device__ __forceinline__ void gdb_sts_gv_if(unsigned addr, gv_t v, bool on) {
#if GDB_GRADIENT
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q st.shared.v2.f32 [%0], {%1, %2}; }" ::"r"(addr), "f"(v.x), "f"(v.y),
                 "r"((unsigned)on)
                 : "memory");
#else
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f32 [%0], %1; }" ::"r"(addr), "f"(v), "r"((unsigned)on)
                 : "memory");
#endif
}
// An address the compiler must keep in a register: without this it re-derives
// the shared-memory layout (~20 uniform instructions) at every single access.
__device__ __forceinline__ unsigned gdb_opaque(unsigned x) {
    unsigned y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint2 gdb_lds_u2(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
// value held by virtual lane `src` (= lane + 32 * slot) of this warp
template<int WPT> __device__ __forceinline__ gv_t gdb_shfl_vlane(const gv_t (&acc)[WPT], unsigned src) {
    gv_t out;
#pragma unroll
    for (int q = 0; q < WPT; ++q) {
#if GDB_GRADIENT
        const gv_t t = make_float2(__shfl_sync(0xffffffffu, acc[q].x, (int)(src & 31u)),
                                   __shfl_sync(0xffffffffu, acc[q].y, (int)(src & 31u)));
#else
        const gv_t t = __shfl_sync(0xffffffffu, acc[q], (int)(src & 31u));
#endif
        if (q == 0 || (src >> 5) == (unsigned)q) out = t;
    }
    return out;
}

// Slots of a column that found no helper lane (the warp ran out of lanes): the
// owner walks them through the row index.  Out of line on purpose -- rare, and
// the caller is replicated per row.
__device__ __noinline__ gv_t gdb_small_overflow(const unsigned *rowadj1, const unsigned *rowptr2, const unsigned *rowadj2,
                                                const unsigned *lanemap2, const unsigned *vovf, const float *W,
                                                const gv_t *pbuf,
                                                unsigned k1beg, unsigned k1end, unsigned wstride, unsigned ovf0, unsigned n2,
                                                unsigned pos, unsigned t0) {
    gv_t acc = gv_make(0.f, 0.f);
    const unsigned col = lanemap2[pos] & 0xffffu;
    const unsigned kbeg = rowptr2[col], deg = rowptr2[col + 1] - kbeg;
    for (unsigned k1 = k1beg; k1 < k1end; ++k1) {
        const unsigned j1 = rowadj1[k1] & 0xffffu;
        for (unsigned t = t0; t < deg; ++t) {
            const float w = W[k1 * wstride + ovf0 + vovf[col] + (t - t0)];
            const unsigned j2 = lanemap2[rowadj2[kbeg + t] & 0xffffu] >> 16;
            acc = gv_fma(w, pbuf[j1 * n2 + j2], acc);
        }
    }
    return acc;
}

struct gdb_small_graph {
    const float *degree;
    const node_t *node;
    const edge_t *edge;
    const unsigned *emeta, *rowptr, *rowadj, *rowpos, *lanemap;
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
    v.rowpos = reinterpret_cast<const unsigned *>(base + h->off_ellslot);
    v.lanemap = reinterpret_cast<const unsigned *>(base + h->off_lanemap);
    v.n = h->n_node;
    v.nnz = h->nnz;
    v.n_tile = h->n_tile;
    return v;
}

#if GDB_BUILD_MASK & 2
extern "C" __global__ void __launch_bounds__(GDB_BLOCK, GDB_MIN_BLOCKS_SMALL)
    mlgk_solve_small(const __grid_constant__ gdb_params P) {
    extern __shared__ __align__(16) unsigned char gdb_smem[];
    __shared__ unsigned s_job[2][2];                     // (ja, jb) of the two pipeline slots
    __shared__ __align__(8) unsigned long long s_bar[2]; // mbarriers: blob bytes have landed
    __shared__ float s_red[2 * 4 * (GDB_WARPS > 0 ? GDB_WARPS : 1)];
    int flip = 0;
    const gdb_params_fixed &F = P.f;

    // Staging pipeline.  The dynamic shared memory starts with TWO blob buffers
    // of F.blob_slot bytes.  Thread 0 claims the next job and issues one TMA
    // bulk copy per graph blob (cp.async.bulk, global -> shared, completion on
    // an mbarrier) into the idle buffer while the CTA solves the current pair,
    // so the global-memory latency of staging is hidden behind the solve.
    // (buffers are addressed as gdb_smem + slot * blob_slot so that the compiler keeps
    // the shared address space and emits LDS, not generic LD)
    unsigned char *const work = gdb_smem + 2 * F.blob_slot;
    if (threadIdx.x == 0) {
        gdb_mbar_init(&s_bar[0], 1);
        gdb_mbar_init(&s_bar[1], 1);
        gdb_fence_mbar_init();
    }
    auto prefetch = [&](int slot) {  // thread 0 only
        const unsigned long long job = atomicAdd(F.counters, 1ull);
        unsigned a = 0xffffffffu, b = 0;
        if (job < F.n_jobs) {
            gdb_decode_job(F, job, a, b);
#if GDB_TMA_STAGE
            const gdb_graph_ref r1 = F.graphs[a], r2 = F.graphs[b];
            const unsigned bytes = r1.bytes + (a == b ? 0u : r2.bytes);
            gdb_fence_proxy_async();  // earlier generic reads of this buffer are done
            gdb_mbar_expect_tx(&s_bar[slot], bytes);
            unsigned char *const dst = gdb_smem + slot * F.blob_slot;
            gdb_bulk_g2s(dst, r1.blob, r1.bytes, &s_bar[slot]);
            if (a != b) gdb_bulk_g2s(dst + r1.bytes, r2.blob, r2.bytes, &s_bar[slot]);
#endif
        }
        s_job[slot][0] = a;
        s_job[slot][1] = b;
    };
    gdb_group_sync();
    if (threadIdx.x == 0) prefetch(0);
    int slot = 0;
    unsigned phase[2] = {0u, 0u};

    while (true) {
        gdb_group_sync();  // previous pair finished everywhere; s_job[slot] is visible
        const unsigned ja = s_job[slot][0], jb = s_job[slot][1];
        if (ja == 0xffffffffu) break;
#if GDB_TMA_STAGE
        if (threadIdx.x == 0) prefetch(slot ^ 1);
#endif
        const bool same = (ja == jb);
#if GDB_TMA_STAGE
        gdb_mbar_wait(&s_bar[slot], phase[slot]);
        phase[slot] ^= 1u;
        unsigned char *const base1 = gdb_smem + slot * F.blob_slot;
        const unsigned char *base2 = same ? base1 : base1 + reinterpret_cast<const gdb_graph_hdr *>(base1)->blob_bytes;
        slot ^= 1;
#else
        // synchronous staging (A/B reference for the TMA pipeline)
        unsigned char *const base1 = gdb_smem;
        const unsigned char *base2 = base1;
        {
            const gdb_graph_ref r1 = F.graphs[ja], r2 = F.graphs[jb];
            gdb_copy16(base1, r1.blob, r1.bytes);
            if (!same) {
                gdb_copy16(base1 + r1.bytes, r2.blob, r2.bytes);
                base2 = base1 + r1.bytes;
            }
            gdb_group_sync();
            if (threadIdx.x == 0) prefetch(0);  // claims the next job only
        }
#endif
        unsigned char *const gdb_smem_work = work;
        // ---- roles.  K and its Jacobian are symmetric in the two graphs, so the roles
        //      are chosen per pair: the columns (lanes) go to the graph whose virtual
        //      columns fit the lanes, then to the larger one (better lane utilisation,
        //      fewer rows per warp).  The output position still follows (ja, jb). --------
#if GDB_ADJ == 2
#define GDB_VCOLS(h) ((h)->vcols & 0xffffu)
#else
#define GDB_VCOLS(h) ((h)->vcols >> 16)
#endif
        constexpr int VL = 32 * GDB_WPT;  // virtual lanes of a warp
#if GDB_NODAL == 0 && !defined(GDB_NO_ROLE_SWAP)
        bool swap_roles;
        {
            const gdb_graph_hdr *ha = reinterpret_cast<const gdb_graph_hdr *>(base1);
            const gdb_graph_hdr *hb = reinterpret_cast<const gdb_graph_hdr *>(base2);
            const int sa = ha->n_node + (GDB_VCOLS(ha) <= (unsigned)VL ? 65536 : 0);
            const int sb = hb->n_node + (GDB_VCOLS(hb) <= (unsigned)VL ? 65536 : 0);
            swap_roles = sa > sb;
        }
        const gdb_small_graph g1 = gdb_small_view(swap_roles ? base2 : base1);
        const gdb_small_graph g2 = gdb_small_view(swap_roles ? base1 : base2);
#else
        const gdb_small_graph g1 = gdb_small_view(base1), g2 = gdb_small_view(base2);
#endif
        const int n1 = g1.n, n2 = g2.n, N = n1 * n2, nnz1 = g1.nnz, nnz2 = g2.nnz;

        // ---- work area: [step table | lane tables | p | W p | W] -----------------------
        //  rtab[i1]  (first, one-past-last) step-table address of row i1 of G1
        //  ktab[k1]  (W row, p row) shared-window addresses of element k1 of G1 (row order):
        //            one 64-bit broadcast load per matvec step
        //  vown[v]   virtual lane v -> column | chunk << 16   (chunk = GDB_ADJ neighbour slots)
        //  vhelp[c]  column c -> first helper lane | helpers << 16
        //  vovf[c]   column c -> index of its first overflow slot (slots without a lane)
        //  wslot[k2] element k2 of G2 (row order) -> float index of its W entry in a W row
        //  vinfo     {lanes in use, most helpers of a column, overflow slots}
        uint2 *ktab = reinterpret_cast<uint2 *>(gdb_smem_work);
        uint2 *rtab = reinterpret_cast<uint2 *>(gdb_smem_work + (((unsigned)nnz1 * 8u + 15u) & ~15u));
        unsigned *vown = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(rtab) + (((unsigned)n1 * 8u + 15u) & ~15u));
        unsigned *vhelp = vown + VL;
        unsigned *vovf = vhelp + ((n2 + 3) & ~3);
        unsigned *wslot = vovf + ((n2 + 3) & ~3);
        unsigned *vinfo = wslot + ((nnz2 + 3) & ~3);
        gv_t *pbuf = reinterpret_cast<gv_t *>(vinfo + 4);
#if GDB_ROLL_ROWS
        gv_t *wpbuf = pbuf + ((N + 3) & ~3);  // W p of the current iteration (hand-off to the row registers)
        float *W = reinterpret_cast<float *>(wpbuf + ((N + 3) & ~3));
#else
        float *W = reinterpret_cast<float *>(pbuf + ((N + 3) & ~3));
        gv_t *wpbuf = reinterpret_cast<gv_t *>(W);  // setup hand-off only: W is filled after it
#endif
        const unsigned W_sa = (unsigned)__cvta_generic_to_shared(W);  // shared-window addresses
        const unsigned p_sa = (unsigned)__cvta_generic_to_shared(pbuf);
        const unsigned ktab_sa = gdb_opaque((unsigned)__cvta_generic_to_shared(ktab));
        const int lane = (int)gdb_opaque(threadIdx.x & 31u);  // kept in a register (S2R is slow)

        // ---- lane tables (warp 0).  Lane position pos < n2 owns the column lanemap[pos]
        //      of G2 (degree-sorted) and its first GDB_ADJ neighbour slots.  A column
        //      with more neighbours gets HELPER lanes (n2, n2 + 1, ...: the lanes a warp
        //      would otherwise idle), one per further chunk of GDB_ADJ slots, highest
        //      degree first; what is left when the lanes run out (rare) is the overflow
        //      that the owner walks through the row index. ------------------------------
        if (threadIdx.x < 32) {
            unsigned next = (unsigned)n2, most = 0u, ovf = 0u;
            for (int p0 = 0; p0 < n2; p0 += 32) {
                const int pos = p0 + lane;
                const bool live = pos < n2;
                const unsigned col = live ? (g2.lanemap[pos] & 0xffffu) : 0u;
                const unsigned deg = live ? g2.rowptr[col + 1] - g2.rowptr[col] : 0u;
                const unsigned want = deg > (unsigned)GDB_ADJ ? (deg - 1u) / (unsigned)GDB_ADJ : 0u;
                unsigned incl = want;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const unsigned start = next + incl - want;
                const unsigned got = start >= (unsigned)VL ? 0u : min(want, (unsigned)VL - start);
                const unsigned covered = (unsigned)GDB_ADJ * (1u + got);
                const unsigned left = deg > covered ? deg - covered : 0u;  // slots without a lane
                unsigned lincl = left;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, lincl, o);
                    if (lane >= o) lincl += t;
                }
                if (live) {
                    vown[pos] = col;
                    vhelp[col] = min(start, 0xffffu) | (got << 16);
                    vovf[col] = ovf + lincl - left;
                    for (unsigned j = 0; j < got; ++j) vown[start + j] = col | ((j + 1u) << 16);
                }
                next += __shfl_sync(0xffffffffu, incl, 31);
                ovf += __shfl_sync(0xffffffffu, lincl, 31);
                most = max(most, got);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {  // ovf is uniform already
                most = max(most, __shfl_xor_sync(0xffffffffu, most, o));
            }
            if (lane == 0) {
                vinfo[0] = min(next, (unsigned)VL);
                vinfo[1] = most;
                vinfo[2] = ovf;
            }
        }

        // ---- per-worker setup: diagonal, rhs, CG start ---------------------------------
        const float Q = 1.0f / (1.0f - F.q), Q2 = Q * Q;
        // rows of G1 are dealt to ALL warps of the CTA in equal blocks of at most GDB_RPW
        // (balanced: 20 rows on 4 warps = 5 + 5 + 5 + 5), so every loop over the elements
        // of a row is warp-uniform
        const int w_rows = (n1 + GDB_WARPS - 1) / GDB_WARPS;
        const int w_row0 = min((int)(threadIdx.x >> 5) * w_rows, n1);
        const int w_row1 = min(w_row0 + w_rows, n1);  // one past this warp's last row
        const int w_nrow = w_row1 - w_row0;
        // p and W are laid out by lane position, the graph data by node
#define GDB_POS(s) (lane + 32 * (s))
#define GDB_LIVE(s) (GDB_POS(s) < n2)
        // this thread's elements of p: row r of slot s at w_psa[s] + r * w_prow, valid for r < w_nown[s]
        const unsigned w_prow = gdb_opaque((unsigned)n2 * (unsigned)sizeof(gv_t));
        const unsigned w_rtsa = gdb_opaque((unsigned)__cvta_generic_to_shared(rtab) + (unsigned)w_row0 * 8u);
        unsigned w_psa[GDB_WPT];
        int w_nown[GDB_WPT];
#pragma unroll
        for (int s = 0; s < GDB_WPT; ++s) {
            w_psa[s] = gdb_opaque(p_sa + (unsigned)(w_row0 * n2 + GDB_POS(s)) * (unsigned)sizeof(gv_t));
            w_nown[s] = (int)gdb_opaque((unsigned)(GDB_LIVE(s) ? w_nrow : 0));
        }
        // The heavy per-element work (node kernel, Jacobians, the matvec) runs in
        // ROLLED loops over the rows and hands its results to / from the row registers
        // through shared memory: the register arrays need compile-time indices, but
        // unrolling the heavy bodies GDB_RPW times overflows the instruction cache
        // (measured: stall_no_instruction 0.2 -> 2.4 per issue at 6800 SASS instructions).
        float diag[GDB_WPT][GDB_RPW];
        gv_t xv[GDB_WPT][GDB_RPW], rv[GDB_WPT][GDB_RPW], apv[GDB_WPT][GDB_RPW];
        float rho[GV_N];
#pragma unroll
        for (int k = 0; k < GV_N; ++k) rho[k] = 0.f;
#pragma unroll
        for (int s = 0; s < GDB_WPT; ++s) {
            const int pos = GDB_POS(s);
            if (GDB_LIVE(s)) {
                const int i2 = (int)(g2.lanemap[pos] & 0xffffu);
                const node_t &u2 = g2.node[i2];
                const float d2 = g2.degree[i2] * Q2;
#if GDB_GRADIENT && !GDB_NGRAD
                const float p2 = P.p_start(u2);
#endif
#pragma unroll 1
                for (int i1 = w_row0; i1 < w_row1; ++i1) {
                    const node_t &u1 = g1.node[i1];
                    const float dx = g1.degree[i1] * d2;
                    // diagonal Dx / Vx in the first float of p, rhs in the W p slot
                    pbuf[i1 * n2 + pos] = gv_make(__fdividef(dx, P.node_kernel(u1, u2)), 0.f);
#if GDB_GRADIENT && !GDB_NGRAD
                    wpbuf[i1 * n2 + pos] = gv_make(dx, P.p_start(u1) * p2);
#else
                    wpbuf[i1 * n2 + pos] = gv_make(dx, 0.f);
#endif
                }
            }
#pragma unroll
            for (int r = 0; r < GDB_RPW; ++r) {
                const int i1 = w_row0 + r;
                diag[s][r] = 1.f;
                xv[s][r] = gv_make(0.f, 0.f);
                rv[s][r] = gv_make(0.f, 0.f);
                apv[s][r] = gv_make(0.f, 0.f);
                if (GDB_LIVE(s) && i1 < w_row1) {
                    const float d = gv_get(pbuf[i1 * n2 + pos], 0);
                    const gv_t ri = wpbuf[i1 * n2 + pos];
                    const gv_t z = gv_scale(__fdividef(1.0f, d), ri);
                    diag[s][r] = d;
                    rv[s][r] = ri;
                    pbuf[i1 * n2 + pos] = z;
#pragma unroll
                    for (int k = 0; k < GV_N; ++k) rho[k] = fmaf(gv_get(ri, k), gv_get(z, k), rho[k]);
                }
            }
        }
        gdb_group_sum_n(rho, s_red, flip);  // barrier: the lane tables are visible

        // ---- virtual lanes of this thread, W layout ------------------------------------
        // W is indexed [k1][v * GDB_ADJ + k]: k1 = element of G1 in row order, (v, k) =
        // k-th slot of virtual lane v, zero where a chunk has fewer neighbours; the
        // overflow elements (if any) follow at [ovf0 + ...].  A lane's slots are one
        // aligned 64/128-bit load, and W rows are visited with a constant stride.
        const unsigned nvl = vinfo[0], w_most = vinfo[1];
        const bool w_ovf = vinfo[2] != 0u;
        const bool w_rare = w_ovf || w_most > 1u;
        const unsigned ovf0 = (nvl * (unsigned)GDB_ADJ + 3u) & ~3u;
        const int wstride = (int)(ovf0 + ((vinfo[2] + 3u) & ~3u));  // floats per W row
        unsigned w_woff[GDB_WPT];           // byte offset of the lane's slots in a W row
        unsigned w_xoff[GDB_WPT][GDB_ADJ];  // byte offsets of the neighbours in a p row
        unsigned w_help[GDB_WPT];           // owners: first helper lane | helpers << 16
#pragma unroll
        for (int s = 0; s < GDB_WPT; ++s) {
            const unsigned v = (unsigned)GDB_POS(s);
            const bool used = v < nvl;
            const unsigned own = used ? vown[v] : 0u;
            const unsigned col = own & 0xffffu, chunk = own >> 16;
            const unsigned kbeg = g2.rowptr[col] + chunk * (unsigned)GDB_ADJ;
            const unsigned kend = used ? min(g2.rowptr[col + 1], kbeg + (unsigned)GDB_ADJ) : kbeg;
            w_woff[s] = used ? v * (unsigned)(GDB_ADJ * 4) : 0u;
#pragma unroll
            for (int k = 0; k < GDB_ADJ; ++k)
                w_xoff[s][k] =
                    (kbeg + k < kend) ? (g2.lanemap[g2.rowadj[kbeg + k] & 0xffffu] >> 16) * (unsigned)sizeof(gv_t) : 0u;
            w_help[s] = (GDB_LIVE(s) && chunk == 0u) ? vhelp[col] : 0u;
        }
        for (int k1 = threadIdx.x; k1 < nnz1; k1 += GDB_BLOCK)
            ktab[k1] = make_uint2(W_sa + (unsigned)k1 * (unsigned)(wstride * 4),
                                  p_sa + (g1.rowadj[k1] & 0xffffu) * ((unsigned)n2 * (unsigned)sizeof(gv_t)));
        for (int i1 = threadIdx.x; i1 < n1; i1 += GDB_BLOCK)
            rtab[i1] = make_uint2(ktab_sa + g1.rowptr[i1] * 8u, ktab_sa + g1.rowptr[i1 + 1] * 8u);
        for (int k2 = threadIdx.x; k2 < nnz2; k2 += GDB_BLOCK) {
            const unsigned rp = g2.rowpos[k2], col = rp & 0xffffu, t = rp >> 16;
            const unsigned chunk = t / (unsigned)GDB_ADJ, slot = t % (unsigned)GDB_ADJ;
            const unsigned hv = vhelp[col];
            unsigned at;
            if (chunk == 0u)
                at = (g2.lanemap[col] >> 16) * (unsigned)GDB_ADJ + slot;
            else if (chunk - 1u < (hv >> 16))
                at = ((hv & 0xffffu) + chunk - 1u) * (unsigned)GDB_ADJ + slot;
            else
                at = ovf0 + vovf[col] + (t - (unsigned)GDB_ADJ * (1u + (hv >> 16)));
            wslot[k2] = at;
        }

        // ---- W = w1 w2 kE(e1, e2), once per pair: zero fill, then one pass over the
        //      nnz1 x nnz2 real element pairs.  The element-pair passes (here and the
        //      edge Jacobian in the epilogue) split the CTA into groups of ep_gs
        //      threads over the elements of G2 (the smallest power-of-two multiple of
        //      32 that divides the block and covers nnz2), the groups share out G1. ---
        unsigned ep_gs = 32u;
        while (ep_gs < (unsigned)nnz2 && ep_gs * 2u <= (unsigned)GDB_BLOCK && (unsigned)GDB_BLOCK % (ep_gs * 2u) == 0u) ep_gs *= 2u;
        const unsigned ep_ng = (unsigned)GDB_BLOCK / ep_gs, ep_grp = threadIdx.x / ep_gs, ep_lane = threadIdx.x % ep_gs;
        {
            float4 *W4 = reinterpret_cast<float4 *>(W);
            for (int idx = threadIdx.x; idx < (nnz1 * wstride) / 4; idx += GDB_BLOCK) W4[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            gdb_group_sync();
            // a thread keeps ONE element of G2 (edge, slot) in registers and walks
            // the elements of G1 that its group is dealt: no index arithmetic and
            // one broadcast load per product
            for (unsigned k2 = ep_lane; k2 < (unsigned)nnz2; k2 += ep_gs) {
                const edge_t e2 = g2.edge[g2.rowadj[k2] >> 16];
                float *Wk = W + wslot[k2];
                for (unsigned k1 = ep_grp; k1 < (unsigned)nnz1; k1 += ep_ng)
                    Wk[k1 * (unsigned)wstride] = gdb_edge_value(P, g1.edge[g1.rowadj[k1] >> 16], e2);
            }
        }
        gdb_group_sync();  // W, the step table and p complete

        // matvec: W p of every owned element, one row of G1 at a time
        auto row_wp = [&](int i1, gv_t (&acc)[GDB_WPT], auto wd) {
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) acc[s] = gv_make(0.f, 0.f);
            const uint2 row = gdb_lds_u2(w_rtsa + (unsigned)(i1 - w_row0) * 8u);  // steps of this row
            GDB_UNROLL(GDB_K1_UNROLL)  // the body is replicated per row already: keep the code in the instruction cache
            for (unsigned ka = row.x; ka != row.y; ka += 8u) {  // warp-uniform trip count
                const uint2 step = gdb_lds_u2(ka);  // (W row, p row)
#pragma unroll
                for (int s = 0; s < GDB_WPT; ++s) {
                    // All lanes load: empty slots hold W = 0 and point at position 0
                    // of the p row, unused lanes read lane 0's slots.
#if GDB_ADJ == 2
                    const float2 w2 = gdb_lds_f2(step.x + w_woff[s] + wd());
                    const gv_t p0 = gdb_lds_gv(step.y + w_xoff[s][0]);
                    const gv_t p1 = gdb_lds_gv(step.y + w_xoff[s][1]);
                    acc[s] = gv_fma(w2.x, p0, acc[s]);
                    acc[s] = gv_fma(w2.y, p1, acc[s]);
#else
                    const float4 w4 = gdb_lds_f4(step.x + w_woff[s] + wd());
                    const gv_t p0 = gdb_lds_gv(step.y + w_xoff[s][0]);
                    const gv_t p1 = gdb_lds_gv(step.y + w_xoff[s][1]);
                    const gv_t p2 = gdb_lds_gv(step.y + w_xoff[s][2]);
                    const gv_t p3 = gdb_lds_gv(step.y + w_xoff[s][3]);
                    acc[s] = gv_fma(w4.x, p0, acc[s]);
                    acc[s] = gv_fma(w4.y, p1, acc[s]);
                    acc[s] = gv_fma(w4.z, p2, acc[s]);
                    acc[s] = gv_fma(w4.w, p3, acc[s]);
#endif
                }
            }
            // helpers hand their partial sums to the owning lane (fixed order).  The first
            // exchange is unconditional -- lanes without a helper add nothing -- so that the
            // common case of molecular graphs (one helper at most, no overflow) is straight-line
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) {
                const gv_t t = gdb_shfl_vlane(acc, w_help[s] & 0xffffu);
                if (w_help[s] >> 16) acc[s] = gv_add(acc[s], t);
            }
            if (w_rare) {  // uniform per pair: more helpers per column, or slots without a lane
#pragma unroll 1
                for (unsigned h = 1; h < w_most; ++h) {
#pragma unroll
                    for (int s = 0; s < GDB_WPT; ++s) {
                        const unsigned src = (w_help[s] & 0xffffu) + h;
                        const gv_t t = gdb_shfl_vlane(acc, src);
                        if (h < (w_help[s] >> 16)) acc[s] = gv_add(acc[s], t);
                    }
                }
                if (w_ovf) {
#pragma unroll
                    for (int s = 0; s < GDB_WPT; ++s)
                        if (GDB_LIVE(s))
                            acc[s] = gv_add(acc[s], gdb_small_overflow(g1.rowadj, g2.rowptr, g2.rowadj, g2.lanemap, vovf, W + wd() / 4u, pbuf,
                                                                       g1.rowptr[i1], g1.rowptr[i1 + 1], (unsigned)wstride, ovf0, (unsigned)n2,
                                                                       (unsigned)GDB_POS(s),
                                                                       (1u + (w_help[s] >> 16)) * (unsigned)GDB_ADJ));
                }
            }
        };

#if GDB_NGRAD
        // ---- nodal Jacobian by forward sensitivities (reference template.cu:226-418 re-solves
        //      twice per hyper-parameter for a central difference): R_i = xs_i p1 p2; d/dp_m is
        //      explicit; for q, node and edge parameters  dx/dt = A^-1 (db/dt - dA/dt x)  is one
        //      more solve with the SAME cached operator W -- two parameters per round on the
        //      two float2 lanes:
        //        q:      rhs = 2Q Dx (1 - x / Vx)
        //        node m: rhs = Dx / Vx^2 dVx_m x          (+ lmin: R -= dVx_m p1 p2)
        //        edge m: rhs = (dW_m) x: dW_m is cached like W in a second buffer and applied
        //                with the matvec of the CG loop
        constexpr int NG_SENS = GDB_NJ - GDB_NP;
        constexpr int NG_ROUNDS = 1 + (NG_SENS + 1) / 2;
        const unsigned w_floats = ((unsigned)(nnz1 * wstride) + 3u) & ~3u;
        float *W2 = W + w_floats;                     // dW_m, same layout as W
        float *Xb = W2 + w_floats;                    // x of the value solve, by node
        gv_t *rhsb = reinterpret_cast<gv_t *>(Xb + ((N + 3) & ~3));  // right-hand sides, by lane position
        const unsigned ngI1 = F.starts[ja] - F.row0, ngI2 = F.starts[jb] - F.col0;
        const unsigned long long ngplane = (unsigned long long)F.nX * F.nY;
        (void)ngI2;
        (void)ngplane;
        auto write_nodal = [&](int i1, int i2, int m, float val) {
#if GDB_NODAL == 2
            F.grad[(unsigned long long)(ngI1 + i1 + i2 * n1) + (unsigned long long)m * F.nX] = val;
#elif GDB_DIAGONAL
            if (i1 == i2) F.grad[(unsigned long long)(ngI1 + i1) + (unsigned long long)m * F.nX] = val;
#else
            F.grad[(unsigned long long)(ngI1 + i1) + (unsigned long long)(ngI2 + i2) * F.nX + m * ngplane] = val;
#if GDB_SYMMETRIC
            if (!same) F.grad[(unsigned long long)(ngI2 + i2) + (unsigned long long)(ngI1 + i1) * F.nX + m * ngplane] = val;
#endif
#endif
        };
        bool w2_clean = false;
        int iters = 0;  // summed over all rounds
#pragma unroll 1
        for (int round = 0; round < NG_ROUNDS; ++round) {
        const int ng_m0 = GDB_NP + 2 * (round - 1);  // Jacobian index of lane 0 of this round (lane 1: + 1)
        if (round > 0) {
            gdb_group_sync();  // the outputs of the previous round have been read from p
#if GDB_NE > 0
            const bool any_edge = ng_m0 + 1 >= GDB_NP + 1 + GDB_NV;  // uniform
            if (any_edge) {  // p := (x, x) by lane position: what the dW matvec gathers
#pragma unroll
                for (int s = 0; s < GDB_WPT; ++s)
                    if (GDB_LIVE(s)) {
                        const int i2 = (int)(g2.lanemap[GDB_POS(s)] & 0xffffu);
#pragma unroll 1
                        for (int i1 = w_row0; i1 < w_row1; ++i1) {
                            const float xi = Xb[i1 * n2 + i2];
                            pbuf[i1 * n2 + GDB_POS(s)] = gv_make(xi, xi);
                        }
                    }
            }
#endif
            // element-wise right-hand sides (q and node parameters); edge parameters start at 0
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s)
                if (GDB_LIVE(s)) {
                    const int i2 = (int)(g2.lanemap[GDB_POS(s)] & 0xffffu);
                    const node_t &u2 = g2.node[i2];
                    const float d2 = g2.degree[i2] * Q2;
#pragma unroll 1
                    for (int i1 = w_row0; i1 < w_row1; ++i1) {
                        const node_t &u1 = g1.node[i1];
                        const float dx = g1.degree[i1] * d2;
                        const float v = P.node_kernel(u1, u2);
                        const float xi = Xb[i1 * n2 + i2];
                        float comp[2] = {0.f, 0.f};
#if GDB_NV > 0
                        float dv[GDB_NV];
                        P.node_kernel.jacobian(u1, u2, dv);
#endif
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int m = ng_m0 + c;
                            if (m == GDB_NP) comp[c] = 2.f * Q * dx * (1.f - __fdividef(xi, v));
#if GDB_NV > 0
                            else if (m < GDB_NP + 1 + GDB_NV) {
                                float dvm = 0.f;
#pragma unroll
                                for (int k = 0; k < GDB_NV; ++k) dvm = (k == m - GDB_NP - 1) ? dv[k] : dvm;
                                comp[c] = __fdividef(dx, v * v) * dvm * xi;
                            }
#endif
                        }
                        rhsb[i1 * n2 + GDB_POS(s)] = gv_make(comp[0], comp[1]);
                    }
                }
#if GDB_NE > 0
            // edge parameters: dW_m into the second buffer (zero pattern as W), then one matvec
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int m = ng_m0 + c;
                if (m < GDB_NP + 1 + GDB_NV || m >= GDB_NJ) continue;  // uniform
                const int ke = m - (GDB_NP + 1 + GDB_NV);
                if (!w2_clean) {
                    float4 *W4 = reinterpret_cast<float4 *>(W2);
                    for (int idx = threadIdx.x; idx < (int)(w_floats / 4u); idx += GDB_BLOCK) W4[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
                    w2_clean = true;
                }
                gdb_group_sync();  // zero fill / the previous use of W2 is over; p = (x, x) and rhsb visible
                for (unsigned k2 = ep_lane; k2 < (unsigned)nnz2; k2 += ep_gs) {
                    const edge_t e2 = g2.edge[g2.rowadj[k2] >> 16];
                    float *Wk = W2 + wslot[k2];
                    for (unsigned k1 = ep_grp; k1 < (unsigned)nnz1; k1 += ep_ng) {
                        const edge_t &e1 = g1.edge[g1.rowadj[k1] >> 16];
                        float de[GDB_NE];
                        P.edge_kernel.jacobian(e1.label, e2.label, de);
                        float dem = 0.f;
#pragma unroll
                        for (int k = 0; k < GDB_NE; ++k) dem = (k == ke) ? de[k] : dem;
#if GDB_WEIGHTED
                        dem *= e1.weight * e2.weight;
#endif
                        Wk[k1 * (unsigned)wstride] = dem;
                    }
                }
                gdb_group_sync();
#pragma unroll 1
                for (int i1 = w_row0; i1 < w_row1; ++i1) {
                    gv_t acc[GDB_WPT];
                    row_wp(i1, acc, gdb_wdelta{w_floats * 4u});
#pragma unroll
                    for (int s = 0; s < GDB_WPT; ++s)
                        if (GDB_LIVE(s)) {
                            gv_t cur = rhsb[i1 * n2 + GDB_POS(s)];
                            if (c == 0) cur.x = acc[s].x; else cur.y = acc[s].y;
                            rhsb[i1 * n2 + GDB_POS(s)] = cur;
                        }
                }
            }
#endif
            gdb_group_sync();  // every read of p = (x, x) is done, rhsb complete
            // row registers of this round: r = rhs, z = r / diag, p = z, x = 0
#pragma unroll
            for (int k = 0; k < GV_N; ++k) rho[k] = 0.f;
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) {
#pragma unroll
                for (int r = 0; r < GDB_RPW; ++r) {
                    const int i1 = w_row0 + r;
                    xv[s][r] = gv_make(0.f, 0.f);
                    rv[s][r] = gv_make(0.f, 0.f);
                    apv[s][r] = gv_make(0.f, 0.f);
                    if (GDB_LIVE(s) && i1 < w_row1) {
                        const gv_t ri = rhsb[i1 * n2 + GDB_POS(s)];
                        const gv_t z = gv_scale(__fdividef(1.0f, diag[s][r]), ri);
                        rv[s][r] = ri;
                        pbuf[i1 * n2 + GDB_POS(s)] = z;
#pragma unroll
                        for (int k = 0; k < GV_N; ++k) rho[k] = fmaf(gv_get(ri, k), gv_get(z, k), rho[k]);
                    }
                }
            }
            gdb_group_sum_n(rho, s_red, flip);  // barrier: p complete
        }
#endif
        // ---- Jacobi-PCG, both systems at once ---------------------------------------
        // stop when sqrt(r.r) < ftol N  <=>  r.r < (ftol N)^2; bit k of `active`: system k runs
        const float thresh2 = (F.ftol * (float)N) * (F.ftol * (float)N);
        unsigned active = 0u;
#pragma unroll
        for (int k = 0; k < GV_N; ++k) active |= (rho[k] != 0.f ? 1u : 0u) << k;
#if !GDB_NGRAD
        int iters = 0;  // summed over the systems that were still active
#endif
        for (int it = 0; it < N; ++it) {
            if (active == 0u) break;
            iters += __popc(active);

            float pAp[GV_N];
#pragma unroll
            for (int k = 0; k < GV_N; ++k) pAp[k] = 0.f;
#if GDB_ROLL_ROWS
            // rolled over the rows (small code); W p reaches the row registers through wpbuf
#pragma unroll 1
            for (int i1 = w_row0; i1 < w_row1; ++i1) {
                gv_t acc[GDB_WPT];
                row_wp(i1, acc, gdb_wzero{});
#pragma unroll
                for (int s = 0; s < GDB_WPT; ++s)
                    if (GDB_LIVE(s)) wpbuf[i1 * n2 + GDB_POS(s)] = acc[s];
            }
#endif
            // A p = diag p - W p and p . A p on the row registers
#pragma unroll
            for (int r = 0; r < GDB_RPW; ++r) {
                const int i1 = w_row0 + r;
                if (i1 < w_row1) {  // warp-uniform
#if !GDB_ROLL_ROWS
                    gv_t acc[GDB_WPT];
                    row_wp(i1, acc, gdb_wzero{});
#endif
#pragma unroll
                    for (int s = 0; s < GDB_WPT; ++s) {
                        // no branch: lanes without an element (idle or helper lanes) get p = 0, and
                        // whatever their A p is, the update of r below multiplies it by zero
                        const bool own = r < w_nown[s];
#if GDB_ROLL_ROWS
                        const gv_t wp = own ? wpbuf[i1 * n2 + GDB_POS(s)] : gv_make(0.f, 0.f);
#else
                        const gv_t wp = acc[s];
#endif
                        const gv_t pv = gdb_lds_gv_if(w_psa[s] + (unsigned)r * w_prow, own);
                        const gv_t av = gv_fma2(gv_make(diag[s][r], diag[s][r]), pv, gv_neg(wp));
                        apv[s][r] = av;
#pragma unroll
                        for (int k = 0; k < GV_N; ++k) pAp[k] = fmaf(gv_get(pv, k), gv_get(av, k), pAp[k]);
                    }
                }
            }
            gdb_group_sum_n(pAp, s_red, flip);  // barrier: every read of p is done
            float alpha[GV_N];
#pragma unroll
            for (int k = 0; k < GV_N; ++k) {
                if (pAp[k] == 0.f) active &= ~(1u << k);
                alpha[k] = (active >> k) & 1u ? __fdividef(rho[k], pAp[k]) : 0.f;
            }
            const gv_t al = gv_make(alpha[0], alpha[GV_N - 1]);
            gv_t nal[GDB_WPT];  // -alpha for the lanes that own elements
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) nal[s] = w_nown[s] > 0 ? gv_neg(al) : gv_make(0.f, 0.f);
            float sums[2 * GV_N];
#pragma unroll
            for (int k = 0; k < 2 * GV_N; ++k) sums[k] = 0.f;
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) {
#pragma unroll
                for (int r = 0; r < GDB_RPW; ++r) {
                    // straight-line code: rows without an element hold r = A p = 0, diag = 1;
                    // lanes without elements use alpha = 0
                    const gv_t ri = gv_fma2(nal[s], apv[s][r], rv[s][r]);
                    rv[s][r] = ri;
                    const float dinv = __fdividef(1.0f, diag[s][r]);
#pragma unroll
                    for (int k = 0; k < GV_N; ++k) {
                        const float rk = gv_get(ri, k);
                        sums[2 * k] = fmaf(rk, rk, sums[2 * k]);
                        sums[2 * k + 1] = fmaf(rk * dinv, rk, sums[2 * k + 1]);
                    }
                }
            }
            gdb_group_sum_n(sums, s_red, flip);
            float beta[GV_N];
#pragma unroll
            for (int k = 0; k < GV_N; ++k) {
                if (sums[2 * k] < thresh2) active &= ~(1u << k);
                const bool on = (active >> k) & 1u;
                beta[k] = on ? __fdividef(sums[2 * k + 1], rho[k]) : 0.f;
                rho[k] = on ? sums[2 * k + 1] : rho[k];
                if (rho[k] == 0.f) active &= ~(1u << k);
            }
            const gv_t be = gv_make(beta[0], beta[GV_N - 1]);
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) {
#pragma unroll
                for (int r = 0; r < GDB_RPW; ++r) {
                    // x += alpha p rides on the read of p that the update of p needs anyway;
                    // only the two memory operations are predicated
                    const bool own = r < w_nown[s];
                    const unsigned pa = w_psa[s] + (unsigned)r * w_prow;
                    const gv_t pv = gdb_lds_gv_if(pa, own);
                    xv[s][r] = gv_fma2(al, pv, xv[s][r]);
                    gdb_sts_gv_if(pa, gv_fma2(be, pv, gv_scale(__fdividef(1.0f, diag[s][r]), rv[s][r])), own);
                }
            }
            gdb_group_sync();  // p complete before the next matvec
        }

#if GDB_NGRAD
        // publish this round's solutions by node (p is free now) and write the outputs
        gdb_group_sync();
        {
            gv_t *ys = pbuf;
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) {
#pragma unroll
                for (int r = 0; r < GDB_RPW; ++r) {
                    const int i1 = w_row0 + r;
                    if (i1 < w_row1 && GDB_LIVE(s)) ys[i1 * n2 + (int)(g2.lanemap[GDB_POS(s)] & 0xffffu)] = xv[s][r];
                }
            }
            gdb_group_sync();
            for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
                const int i1 = i / n2, i2 = i - i1 * n2;
                const node_t &u1 = g1.node[i1];
                const node_t &u2 = g2.node[i2];
                const float p1 = P.p_start(u1), p2 = P.p_start(u2);
#if GDB_NODAL == 2 || GDB_SYMMETRIC
                const bool sym = same;  // self pair: bit-exact symmetry of the output
#else
                const bool sym = false;
#endif
                if (round == 0) {
                    // the value solve: nodal Gram entries (as the epilogue of the other programs),
                    // x kept for the sensitivity rounds, explicit d/dp planes
                    const float xraw = gv_get(ys[i], 0);
                    Xb[i] = xraw;
                    float xi = sym ? 0.5f * (xraw + gv_get(ys[i2 * n2 + i1], 0)) : xraw;
#if GDB_LMIN == 1
                    xi -= P.node_kernel(u1, u2);
#endif
#if GDB_NODAL == 2
                    F.gram[ngI1 + i1 + i2 * n1] = xi * p1 * p2;
#elif GDB_DIAGONAL
                    if (i1 == i2) F.gram[ngI1 + i1] = xi * p1 * p2;
#else
                    F.gram[(unsigned long long)(ngI1 + i1) + (unsigned long long)(ngI2 + i2) * F.nX] = xi * p1 * p2;
#if GDB_SYMMETRIC
                    if (!same) F.gram[(unsigned long long)(ngI2 + i2) + (unsigned long long)(ngI1 + i1) * F.nX] = xi * p1 * p2;
#endif
#endif
#if GDB_NP > 0
                    float d1[GDB_NP], d2[GDB_NP];
                    P.p_start.jacobian(u1, d1);
                    P.p_start.jacobian(u2, d2);
#pragma unroll
                    for (int m = 0; m < GDB_NP; ++m)
                        write_nodal(i1, i2, m, xi * __fadd_rn(__fmul_rn(d1[m], p2), __fmul_rn(p1, d2[m])));  // no FMA: symmetric under swap
#endif
                } else {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int m = ng_m0 + c;
                        if (m >= GDB_NJ) continue;
                        float val = gv_get(ys[i], c);
                        if (sym) val = 0.5f * (val + gv_get(ys[i2 * n2 + i1], c));
#if GDB_LMIN == 1 && GDB_NV > 0
                        if (m > GDB_NP && m < GDB_NP + 1 + GDB_NV) {
                            float dv[GDB_NV];
                            P.node_kernel.jacobian(u1, u2, dv);
#pragma unroll
                            for (int k = 0; k < GDB_NV; ++k) val -= (k == m - GDB_NP - 1) ? dv[k] : 0.f;
                        }
#endif
                        write_nodal(i1, i2, m, val * p1 * p2);
                    }
                }
            }
        }
        }  // rounds
#endif

        if (threadIdx.x == 0) {
            atomicAdd(F.counters + 1, (unsigned long long)iters);
            atomicAdd(F.counters + 2, (unsigned long long)iters * (unsigned long long)nnz1 * (unsigned long long)nnz2);
            atomicAdd(F.counters + 3, (unsigned long long)iters * (unsigned long long)N);
        }

#if !GDB_NGRAD
        const unsigned I1 = F.starts[ja] - F.row0, I2 = F.starts[jb] - F.col0;
        const unsigned long long plane = (unsigned long long)F.nX * F.nY;
        (void)plane;
        (void)I2;

        // publish x (both systems) in shared memory, indexed by node: the epilogues run
        // rolled loops, and the nodal outputs and the edge-Jacobian pass read elements
        // owned by other threads
        gv_t *xs = pbuf;
#pragma unroll
        for (int s = 0; s < GDB_WPT; ++s) {
#pragma unroll
            for (int r = 0; r < GDB_RPW; ++r) {
                const int i1 = w_row0 + r;
                if (i1 < w_row1 && GDB_LIVE(s)) xs[i1 * n2 + (int)(g2.lanemap[GDB_POS(s)] & 0xffffu)] = xv[s][r];
            }
        }
        gdb_group_sync();

        // ---- epilogue: starting probabilities, Gram entry ----------------------------
#if GDB_NODAL == 2
        for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            float xi = 0.5f * (gv_get(xs[i], 0) + gv_get(xs[i2 * n2 + i1], 0));  // self pair: bit-exact symmetry
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i2]);
#endif
            F.gram[I1 + i1 + i2 * n1] = xi * P.p_start(g1.node[i1]) * P.p_start(g2.node[i2]);
        }
#elif GDB_NODAL == 1 && GDB_DIAGONAL
        for (int i1 = threadIdx.x; i1 < n1; i1 += GDB_BLOCK) {
            float xi = gv_get(xs[i1 * n2 + i1], 0);
#if GDB_LMIN == 1
            xi -= P.node_kernel(g1.node[i1], g2.node[i1]);
#endif
            const float ps = P.p_start(g1.node[i1]);
            F.gram[I1 + i1] = xi * ps * ps;
        }
#elif GDB_NODAL == 1
        for (int i = threadIdx.x; i < N; i += GDB_BLOCK) {
            const int i1 = i / n2, i2 = i - i1 * n2;
            float xi = gv_get(xs[i], 0);
#if GDB_SYMMETRIC
            if (same) xi = 0.5f * (xi + gv_get(xs[i2 * n2 + i1], 0));  // bit-exact symmetry of self pairs
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
        // graph level: K = sum xs p1 p2; with gradients also the node-side Jacobian
        // terms, every worker over its own elements (rolled over the rows)
        {
            constexpr int NACC = 1 + (GDB_GRADIENT ? GDB_NP + 1 + GDB_NV : 0);
            float acc[NACC];
#pragma unroll
            for (int m = 0; m < NACC; ++m) acc[m] = 0.f;
#pragma unroll
            for (int s = 0; s < GDB_WPT; ++s) {
                if (GDB_LIVE(s)) {
                    const int i2 = (int)(g2.lanemap[GDB_POS(s)] & 0xffffu);
                    const node_t &u2 = g2.node[i2];
                    const float p2 = P.p_start(u2);
#if GDB_GRADIENT
                    const float d2 = g2.degree[i2] * Q2;
#endif
#pragma unroll 1
                    for (int i1 = w_row0; i1 < w_row1; ++i1) {
                        const node_t &u1 = g1.node[i1];
                        const float p1 = P.p_start(u1);
                        const gv_t xy = xs[i1 * n2 + i2];
                        const float xi = gv_get(xy, 0);
                        float xsft = xi;
#if GDB_LMIN == 1 || GDB_GRADIENT
                        const float v = P.node_kernel(u1, u2);
#endif
#if GDB_LMIN == 1
                        xsft -= v;
#endif
                        acc[0] = fmaf(xsft, p1 * p2, acc[0]);
#if GDB_GRADIENT
                        // dK/dp_m  = sum (dp1 p2 + p1 dp2) xs
                        // dK/dq    = sum y (2Q Dx) (1 - x / Vx)
                        // dK/dtv_m = sum y x Dx / Vx^2 dVx  [- p1 p2 dVx if lmin]
                        const float yi = gv_get(xy, 1);
                        const float dx = g1.degree[i1] * d2;
#if GDB_NP > 0
                        {
                            float d1[GDB_NP], dd2[GDB_NP];
                            P.p_start.jacobian(u1, d1);
                            P.p_start.jacobian(u2, dd2);
#pragma unroll
                            for (int m = 0; m < GDB_NP; ++m)
                                acc[1 + m] = fmaf(fmaf(d1[m], p2, p1 * dd2[m]), xsft, acc[1 + m]);
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
                }
            }
#if GDB_GRADIENT && GDB_NE > 0
            // dK/dte_m = sum_{i,j} y_i x_j w1 w2 dkE_m(e1, e2): one balanced pass over
            // all element pairs
            float eacc[GDB_NE];
#pragma unroll
            for (int m = 0; m < GDB_NE; ++m) eacc[m] = 0.f;
            for (unsigned e2 = ep_lane; e2 < (unsigned)nnz2; e2 += ep_gs) {  // element of G2 in registers
                const unsigned m2 = g2.emeta[e2];
                const edge_t b = g2.edge[e2];
                const gv_t *xs_y = xs + (m2 & 0xffffu), *xs_x = xs + (m2 >> 16);
                for (unsigned e1 = ep_grp; e1 < (unsigned)nnz1; e1 += ep_ng) {
                    const unsigned m1 = g1.emeta[e1];
                    const float yi = gv_get(xs_y[(m1 & 0xffffu) * n2], 1);
                    float xj = gv_get(xs_x[(m1 >> 16) * n2], 0);
                    const edge_t &a = g1.edge[e1];
#if GDB_WEIGHTED
                    xj *= a.weight * b.weight;
#endif
                    float de[GDB_NE];
                    P.edge_kernel.jacobian(a.label, b.label, de);
#pragma unroll
                    for (int m = 0; m < GDB_NE; ++m) eacc[m] = fmaf(de[m] * yi, xj, eacc[m]);
                }
            }
#endif
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
#if !GDB_DIAGONAL
                const float norm_rs = gdb_norm_scale(F, ja, jb);
                acc[0] *= norm_rs;
#endif
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
                    val = gdb_norm_grad(F, ja, jb, m, norm_rs, acc[0], val);
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
#endif  // !GDB_NGRAD
    }
}
#undef GDB_POS
#undef GDB_LIVE
#undef GDB_VCOLS
#endif  // GDB_BUILD_MASK & 2
