// gdb_internal.h -- definitions shared by the host translation units.
#pragma once
#include <cstdint>

#include "../../include/graphdot_b200.h"

#define GDB_HDR_BYTES 96

// Host mirrors of the device structs in mlgk_solver.cuh.
struct gdb_graph_hdr_host {
    int32_t n_node, n_octile, nnz, n_tile;
    uint32_t off_degree, off_node, off_octile, off_tilerow;
    uint32_t off_edge, off_pool, blob_bytes, flags;
    uint32_t off_emeta, off_rowptr, off_rowadj, off_tileelem;
    uint32_t max_degree;   // largest number of stored elements in a row
    uint32_t off_ellslot;  // u32[nnz] "rowpos": CSR position -> row | (index within the row << 16)
    uint32_t off_lanemap;  // u32[n_node]: [p] & 0xffff = node at degree-sorted position p, [i] >> 16 = position of node i
    uint32_t vcols;        // virtual columns: sum ceil(deg / 2) | sum ceil(deg / 4) << 16 (each >= 1 per node)
    // neighbour-row lists of the large-pair kernel: for every tile row (8 rows) the
    // sorted distinct columns its elements touch -- the rows of the search direction
    // that the matvec of that tile row stages in shared memory
    uint32_t off_tcptr;    // u32[n_tile + 1]: CSR over tile rows into tccol
    uint32_t off_tccol;    // u16[]: distinct columns of a tile row, ascending
    uint32_t off_tcslot;   // u16[nnz]: CSR element k -> index of its column in its tile row's list
    uint32_t max_tc;       // longest list
};
static_assert(sizeof(gdb_graph_hdr_host) == GDB_HDR_BYTES, "header layout");

struct gdb_octile_host {
    uint64_t mask;
    uint32_t start;
    uint16_t trow, tcol;
};
static_assert(sizeof(gdb_octile_host) == 16, "octile layout");

struct gdb_graph_ref_host {
    uint64_t blob;
    uint32_t bytes, n_node;
};
static_assert(sizeof(gdb_graph_ref_host) == 16, "graph ref layout");

struct gdb_params_fixed_host {
    uint64_t graphs, jobs, starts, gram, grad, scratch, counters;
    uint64_t scratch_stride, n_jobs;
    uint32_t job_mode, i0, i1, j0, j1, nX, nY, nJ;
    float q, eps, ftol, gtol;
    uint32_t smem_bytes, row0, col0, norm_n;
    uint64_t norm_diag, norm_ddiag;
    uint32_t blob_slot, pad3;
};
static_assert(sizeof(gdb_params_fixed_host) == 160, "params layout");

// Records `msg` as the calling thread's last error and returns `code`.
int gdb_fail(int code, const char *fmt, ...);
