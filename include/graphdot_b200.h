/* graphdot_b200.h -- C ABI of the B200-native marginalized graph kernel engine.
 *
 * This is the drop-in boundary for the reference's CUDA back end
 * (reference graphdot/kernel/marginalized/_backend_cuda.py).  Plain pointers
 * and sizes only; no torch / pycuda types.  Every entry point names the
 * reference interface it replaces.  All functions return GDB_OK (0) or a
 * negative status; gdb_last_error() then holds a message for the calling
 * thread.  The Python wrapper (graphdot_b200/kernel/marginalized/
 * _backend_b200.py) turns statuses into exceptions and releases the GIL
 * around every call, so one host thread per GPU can drive a job queue.
 */
#ifndef GRAPHDOT_B200_H_
#define GRAPHDOT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDB_OK 0
#define GDB_ERR_INVALID (-1) /* bad argument                                */
#define GDB_ERR_CUDA (-2)    /* CUDA driver/runtime failure, no device      */
#define GDB_ERR_COMPILE (-3) /* NVRTC rejected the program (see log)        */
#define GDB_ERR_NOMEM (-4)
#define GDB_ERR_LAYOUT (-5)  /* host/device struct layout disagreement      */

typedef struct gdb_context_s *gdb_context_t;
typedef struct gdb_program_s *gdb_program_t;
typedef struct gdb_graphset_s *gdb_graphset_t;

/* ---- library ---------------------------------------------------------- */
const char *gdb_version(void);
const char *gdb_last_error(void);
/* The fixed hand-written solver template the microkernel expressions are
 * spliced into (replaces reference kernel/marginalized/template.cu +
 * graphdot/cpp/ *.h as nvcc input). */
const char *gdb_solver_template(void);

/* ---- device context ----------------------------------------------------
 * Replaces pycuda.autoinit / graphdot.cuda.defctx (reference
 * graphdot/cuda/__init__.py:3-7, _backend_cuda.py:49-52).  Uses the device's
 * primary context, so it shares memory and streams with the CUDA runtime
 * (PyTorch). */
typedef struct gdb_device_info {
    int32_t device;
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int32_t max_smem_per_block_optin;
    int32_t max_smem_per_sm;
    int32_t clock_khz;
    int32_t l2_bytes;
    uint64_t total_mem;
    char name[64];
} gdb_device_info;

int gdb_context_create(int device, gdb_context_t *out);
int gdb_context_destroy(gdb_context_t ctx);
int gdb_context_info(gdb_context_t ctx, gdb_device_info *out);
int gdb_context_synchronize(gdb_context_t ctx);

/* Page-locked host buffers for jobs / starts / outputs.  Replaces the
 * managed-memory allocators Backend.array/zeros/empty (reference
 * graphdot/cuda/array.py:14-31, _backend_cuda.py:37-47).  Usable without a
 * context; falls back to nothing: fails if no CUDA device is present. */
int gdb_host_alloc(size_t bytes, void **out);
int gdb_host_free(void *ptr);
/* Page-lock / unlock memory the caller owns (e.g. a POSIX shared-memory
 * mapping that several per-GPU processes fill with Gram tiles: the
 * "gathering Gram tiles back to host" of a multi-GPU run needs no
 * collective).  Portable across contexts. */
int gdb_host_register(void *ptr, size_t bytes);
int gdb_host_unregister(void *ptr);

/* ---- program: NVRTC-compiled solver ------------------------------------
 * Replaces CUDABackend.gencode_kernel / gencode_probability / template
 * rendering / pycuda SourceModule (reference _backend_cuda.py:118-134,
 * :157-228, :282-293).  Inputs are exactly the microkernels' gen_expr()
 * strings and the C++ member declarations of the attribute structs. */
typedef struct gdb_functor_src {
    const char *theta_decl;  /* members of the hyper-parameter struct, e.g.
                                "struct{float32 h;}element;"               */
    uint32_t theta_size;     /* sizeof(np.dtype(kernel.dtype)); may be 0   */
    const char *expr;        /* value expression over x1,x2 (or n)         */
    uint32_t n_jac;
    const char *const *jac;  /* one Jacobian expression per hyper-param    */
} gdb_functor_src;

#define GDB_NODAL_NONE 0
#define GDB_NODAL_FULL 1
#define GDB_NODAL_BLOCK 2

typedef struct gdb_program_desc {
    const char *node_decl;   /* members of node_t                          */
    uint32_t node_size;
    const char *edge_decl;   /* members of the edge label struct           */
    uint32_t edge_label_size, edge_label_align;
    int32_t weighted;        /* edge_t = {float32 weight; label}           */
    gdb_functor_src node_kernel, edge_kernel, p_start;
    /* traits (reference _kernel.py:48-58) -- compile-time specialisation  */
    int32_t diagonal, symmetric, nodal, lmin, eval_gradient;
    int32_t block_size;      /* threads cooperating on one pair; 0 = auto  */
    int32_t workers_per_thread; /* small-pair kernel: columns per lane, 1..4;
                                0 = 1.  That kernel maps one warp to each
                                8-row tile of the first graph and lanes to the
                                columns (nodes of the second graph); it is used
                                when block_size >= 32 * tile rows and
                                32 * workers_per_thread >= nodes             */
    int32_t rows_per_warp;   /* small-pair kernel: rows of the first graph
                                per warp, 1..8; 0 = 8.  Sizes the register
                                arrays; the kernel is used when
                                rows_per_warp * block_size / 32 >= nodes    */
    int32_t slots_per_lane;  /* small-pair kernel: neighbour slots of a
                                column that one lane gathers per step, 2 or
                                4; 0 = 4.  Columns of higher degree borrow
                                the idle lanes of the warp as helpers       */
    const char *extra_options; /* extra NVRTC options, space separated     */
    /* large-pair kernel (one thread-block cluster per pair; used when the CG
     * vectors of the largest pair do not fit in shared memory)            */
    int32_t cluster_size;    /* CTAs per pair: 1, 2, 4 or 8; 0 = 2          */
    int32_t cols_per_lane;   /* columns of the second graph per lane, 1..32;
                                32 * cols_per_lane >= nodes; 0 = 16          */
    int32_t ell_slots;       /* neighbours per column of the second graph
                                held in shared memory (ELL), 1..64; 0 = 12   */
} gdb_program_desc;

typedef struct gdb_program_info {
    int32_t block_size;
    int32_t num_regs;        /* general kernel (mlgk_solve)                */
    int32_t num_regs_small;  /* shared-memory kernel (mlgk_solve_small)    */
    int32_t num_regs_large;  /* cluster kernel (mlgk_solve_large); 0 = none, or not
                              * compiled yet: its NVRTC module is built when a
                              * graph set first needs it */
    int32_t static_smem;
    int32_t local_bytes;     /* spill / stack per thread                   */
    int32_t max_dynamic_smem;
    int32_t n_jac;           /* n_p + 1 + n_node + n_edge                  */
    int32_t from_cache;
    float compile_ms;
} gdb_program_info;

int gdb_program_create(gdb_context_t ctx, const gdb_program_desc *desc,
                       gdb_program_t *out);
int gdb_program_info_get(gdb_program_t prog, gdb_program_info *out);
const char *gdb_program_log(gdb_program_t prog);
const char *gdb_program_source(gdb_program_t prog);
int gdb_program_destroy(gdb_program_t prog);
/* Render the full translation unit without a device (build checks). */
int gdb_render_source(const gdb_program_desc *desc, char **out_malloced);
/* NVRTC-compile for sm_100a without a device or context: reports the cubin
 * size; the compiler log is kept in gdb_last_error() on failure. */
int gdb_program_compile_only(const gdb_program_desc *desc,
                             uint64_t *cubin_bytes);
void gdb_free(void *p);

/* ---- graphs: octile packing and upload ---------------------------------
 * Replaces OctileGraph (reference _octilegraph.py:11-189) and graph_t
 * (reference graphdot/cpp/graph.h:8-33).  A packed graph is ONE
 * position-independent, 16-byte aligned blob
 *   [header 96 B | degree f32[n] | node_t[n] | octile[n_oct]
 *    | tile_row u32[T+1] | edge_t[nnz]
 *    | elem_meta u32[nnz] | row_ptr u32[n+1] | row_adj u32[nnz]
 *    | tile_elem u32[T+1] | row_pos u32[nnz] | lane_map u32[n]
 *    | tc_ptr u32[T+1] | tc_col u16[] | tc_slot u16[nnz]
 *    | variable-length feature pool]
 * with 8x8 octiles sorted by (tile row, tile column), a row-major 64-bit
 * non-zero mask per octile and compact row-major elements.  Degrees are the
 * sums of incident weights (self loops once), 0 replaced by 1 (reference
 * _octilegraph.py:113-139).  elem_meta .. lane_map are a pair-independent row
 * index derived from the octiles at pack time (CSR over rows, degree-sorted
 * lane map; layout in csrc/gdb_internal.h and DESIGN.md section 3), so that
 * the solver kernels never decode bit masks in their hot loops. */
typedef struct gdb_graph_src {
    uint32_t n_node;
    uint32_t n_edge;          /* undirected edges                          */
    const void *nodes;        /* node_t[n_node] in node-index order        */
    const uint32_t *edge_i, *edge_j;
    const float *edge_w;      /* NULL when unweighted                      */
    const void *edge_labels;  /* label struct per undirected edge          */
    const void *pool;         /* variable-length feature data              */
    uint32_t pool_bytes;
} gdb_graph_src;

typedef struct gdb_layout {
    uint32_t node_size;
    uint32_t edge_label_size, edge_label_align;
    int32_t weighted;
    /* byte offsets (inside node_t / edge label) of the 8-byte data-pointer
     * slots of frozen_array members; the host passes pool-relative offsets
     * there and the packer/uploader relocates them. */
    uint32_t n_node_ptr, node_ptr_offset[8];
    uint32_t n_edge_ptr, edge_ptr_offset[8];
} gdb_layout;

/* Size of / write the packed blob of one graph (pure host code). */
int gdb_graph_packed_size(const gdb_layout *layout, const gdb_graph_src *g,
                          uint64_t *bytes);
int gdb_graph_pack(const gdb_layout *layout, const gdb_graph_src *g,
                   void *blob, uint64_t capacity);

/* Pack k graphs at once on `n_threads` host threads (0 = hardware
 * concurrency) from column-concatenated inputs: graph g owns nodes
 * [node_off[g], node_off[g+1]) of `nodes`, undirected edges
 * [edge_off[g], edge_off[g+1]) of edge_i/edge_j/edge_w/edge_labels (end
 * points local to the graph) and bytes [pool_off[g], pool_off[g+1]) of
 * `pool` (pool_off may be NULL: no variable-length features).  Two calls:
 * with blobs == NULL only blob_off[0..k] (prefix sums of the blob sizes) is
 * written; with blobs != NULL of capacity >= blob_off[k] the blobs are
 * written back to back.  Replaces the per-graph Python loop over OctileGraph
 * (reference _backend_cuda.py:252-259, _octilegraph.py:37-177), which is
 * milliseconds per graph. */
typedef struct gdb_batch_src {
    uint32_t n_graphs;
    const uint64_t *node_off, *edge_off, *pool_off;
    const void *nodes;
    const uint32_t *edge_i, *edge_j;
    const float *edge_w;
    const void *edge_labels;
    const void *pool;
} gdb_batch_src;
int gdb_graphs_pack_batch(const gdb_layout *layout, const gdb_batch_src *src,
                          uint64_t *blob_off, void *blobs, uint64_t capacity,
                          int32_t n_threads);

/* Node reordering (pure host): a permutation for Graph.permute -- new index
 * of old node perm[k] is k -- that shrinks the tile footprint of a graph.
 * GDB_REORDER_RCM: reverse Cuthill-McKee (replaces reference
 * graphdot/graph/reorder/rcm.py:7-22).  GDB_REORDER_TILES: greedy growth of
 * 8-node blocks, the role of the reference's partition-based reordering
 * (graphdot/graph/reorder/pbr, which minimises non-empty 8 x 8 tiles with a
 * hypergraph partitioner, pbr/mnom.py:11-24).  gdb_graph_count_tiles counts
 * the non-empty 8 x 8 tiles of the symmetric adjacency, optionally after a
 * permutation (perm may be NULL). */
#define GDB_REORDER_RCM 0
#define GDB_REORDER_TILES 1
int gdb_graph_reorder(uint32_t n_node, uint32_t n_edge, const uint32_t *edge_i,
                      const uint32_t *edge_j, int32_t method, uint32_t *perm);
int gdb_graph_count_tiles(uint32_t n_node, uint32_t n_edge,
                          const uint32_t *edge_i, const uint32_t *edge_j,
                          const uint32_t *perm, uint64_t *n_tiles);

/* Assemble packed blobs into one device-resident graph set. */
int gdb_graphset_create(gdb_context_t ctx, const gdb_layout *layout,
                        uint32_t n_graphs, const void *const *blobs,
                        const uint64_t *blob_bytes, gdb_graphset_t *out);
/* Re-send the staged (pinned) host image to the device: the host->device
 * leg of an end-to-end call. */
int gdb_graphset_upload(gdb_graphset_t gs);
int gdb_graphset_bytes(gdb_graphset_t gs, uint64_t *bytes);
int gdb_graphset_destroy(gdb_graphset_t gs);

/* ---- solve --------------------------------------------------------------
 * Replaces the launch of graph_kernel_solver (reference
 * _backend_cuda.py:303-367, template.cu:29-475): for every job (i, j) solve
 * the product-graph system of graphs i and j and write the Gram entry (and
 * Jacobian) at starts[i], starts[j] of the Fortran-ordered outputs. */
#define GDB_JOBS_LIST 0 /* explicit (i, j) pairs                          */
#define GDB_JOBS_RECT 1 /* all (i, j) with i in [i0,i1), j in [j0,j1)     */
#define GDB_JOBS_TRIU 2 /* all (i, j) with i in [i0,i1), j in [i,j1)      */
#define GDB_OUT_NONE 0
#define GDB_OUT_F64 1
#define GDB_OUT_F32 2

typedef struct gdb_solve_args {
    int32_t job_mode;
    const uint32_t *jobs;     /* host, 2*n_jobs (GDB_JOBS_LIST)            */
    uint64_t n_jobs;
    uint32_t i0, i1, j0, j1;  /* GDB_JOBS_RECT / TRIU                      */
    const uint32_t *starts;   /* host, one per graph (+1)                  */
    uint32_t n_starts;
    float q, eps, ftol, gtol;
    const void *node_theta, *edge_theta, *p_theta; /* raw struct bytes     */
    float *gramian;           /* host; nX*nY floats, Fortran order         */
    float *gradient;          /* host or NULL; nX*nY*nJ floats             */
    uint32_t nX, nY, nJ;
    uint32_t row0, col0;      /* subtracted from starts[i] / starts[j]: lets
                                 a tile of a larger Gram use a tile-sized
                                 output                                    */
    int32_t store_diag;       /* 1 (diagonal programs): keep this solve's
                                 self-similarities (and Jacobians) on the
                                 device for later `normalize` solves over the
                                 same graph set; jobs must be (i, i) for every
                                 graph with starts[i] = i                   */
    int32_t normalize;        /* 1 (graph-level, non-diagonal programs):
                                 write K_ij / sqrt(K_ii K_jj) and its Jacobian
                                 (reference kernel/fix.py:46-73) using the
                                 stored self-similarities                   */
    int32_t upload_graphs;    /* 1: re-send the graph set's pinned host image
                                 first (host->device leg of an end-to-end
                                 call)                                     */
    void *stream;             /* CUstream to run on; NULL = context stream */
    int32_t keep_on_device;   /* 1: skip the device->host copy; outputs
                                 stay in the context's device buffers      */
    /* caller-owned DEVICE outputs (same Fortran layout, nX*nY [*nJ] floats):
     * the kernels write there instead of the context's buffers and nothing
     * is zero-filled first.  With host buffers given as well, finished
     * column blocks are copied from there to the host.                     */
    float *gramian_dev, *gradient_dev;
    /* Pipelined execution: split a TRIU job grid of a symmetric program into
     * launches of `tile` rows, a RECT grid into launches of `tile` columns
     * (0 = one launch).  A finished launch completes a block of output
     * COLUMNS, whose device->host copy (on a second stream) and host-side
     * collection overlap the next launch.                                  */
    uint32_t tile;
    float tile_shrink;        /* 0 or 1: all launches `tile` wide; 0 < s < 1:
                                 every launch s times the previous one (at
                                 least 64): big launches first, so that the
                                 copy and collection of the LAST, exposed
                                 column block are short                    */
    /* Host-side collection, fused with the copy-back (replaces the
     * reshape / active-theta masking / astype of reference
     * _kernel.py:247-264, which is a serial numpy pass over 4(1+nJ) B per
     * pair): out_dtype GDB_OUT_F64 / GDB_OUT_F32 converts every finished
     * column block on the context's host threads into out_gram (nX*nY) and
     * out_grad (nX*nY*n_active, only planes with plane_mask[k] != 0; NULL =
     * all).  `gramian` / `gradient` are then the page-locked float staging. */
    int32_t out_dtype;
    void *out_gram, *out_grad;
    const uint8_t *plane_mask;
    /* 1: return without waiting for the device; the copies into `gramian` /
     * `gradient` are complete after gdb_context_synchronize().  Not combined
     * with out_dtype; diagnostics below are not filled.                   */
    int32_t async;
    /* diagnostics (out) */
    float kernel_ms;          /* device time of the solver kernel          */
    float h2d_ms, d2h_ms;
    uint64_t cg_iterations;   /* total PCG iterations over all jobs        */
    uint64_t matvec_products; /* total nnz1*nnz2 products evaluated        */
    uint64_t vector_elements; /* total N = n1*n2 summed over CG iterations */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t n_launches;
    int32_t used_small_kernel; /* 1: mlgk_solve_small ran, 2: mlgk_solve_large,
                                  0: mlgk_solve                              */
    uint32_t grid, smem_bytes; /* launch configuration used                 */
} gdb_solve_args;

int gdb_solve(gdb_context_t ctx, gdb_program_t prog, gdb_graphset_t gs,
              gdb_solve_args *args);
/* Device pointers of the most recent outputs (keep_on_device). */
int gdb_last_outputs(gdb_context_t ctx, void **gramian_dev,
                     void **gradient_dev);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHDOT_B200_H_ */
