/*
 * pfs_b200.h -- C ABI of libpfs_b200.so: the B200 (sm_100a) message-passing layer.
 *
 * The reference (joshua-lintropic/pfs-neural-net) has no FFI: its boundary for this path is the
 * nn.Module surface of src/gnn.py.  The Python package `pfs-neural-net_b200/gnn.py` mirrors that
 * surface and lowers every module `forward`/`backward` onto the entry points below (ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add).  Each entry point names the
 * reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller allocates every output and the workspace (size from pfs_workspace_bytes);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden streams, no
 *     device synchronisation, no host allocation inside a call;
 *   - return value 0 = ok, negative = error; pfs_last_error() gives the message (thread local);
 *   - nothing throws across the ABI; there is no CPU fallback: without a CUDA device every
 *     compute entry point returns PFS_ERR_CUDA;
 *   - tensors are row-major and contiguous, fp32; a leading "graph" dimension G batches
 *     independent graphs that share one topology and one set of weights: x_s [G,S,F],
 *     x_t [G,T,F], x_e [G,E,F], u [G,F].  BatchNorm statistics are per graph (the reference
 *     processes one graph per step), running buffers are updated graph by graph in order, and
 *     parameter gradients are summed over the G graphs in a fixed order (deterministic).
 */
#ifndef PFS_B200_H
#define PFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFS_ABI_VERSION 1

enum {
    PFS_OK = 0,
    PFS_ERR_ARG = -1,        /* bad argument (null pointer, unsupported F, ...) */
    PFS_ERR_CUDA = -2,       /* CUDA runtime error (message has the cudaGetErrorString) */
    PFS_ERR_UNSUPPORTED = -3,/* valid request this build cannot serve (e.g. fibre degree > tile) */
    PFS_ERR_WORKSPACE = -4   /* workspace too small */
};

enum { PFS_LAYOUT_DENSE = 0, PFS_LAYOUT_CSR = 1 };

/* edges processed by one CTA tile; a fibre's edges never straddle tiles (max fibre degree) */
#define PFS_TILE_EDGES 256

/*
 * Topology of the bipartite graph (reference: `edge_index` [2,E] int64, src/gnn.py:98,135,187).
 *   PFS_LAYOUT_DENSE: the canonical complete graph of src/train.py:94, edge e = k*T + i
 *                     (src = e / T, tgt = e % T); no index arrays are read.
 *   PFS_LAYOUT_CSR:   any edge list.  Arrays come from pfs_build_topology (int32, device):
 *                     edges are visited in fibre-sorted ("CSR") order q = 0..E-1.
 */
typedef struct pfs_topology {
    int32_t layout;
    int32_t G;                 /* graphs in the batch (>= 1) */
    int32_t F;                 /* feature width Fdim */
    int32_t S, T, E;           /* fibres, classes, edges (per graph) */
    const int32_t* csr_rowptr; /* [S+1]  first CSR position of each fibre                     */
    const int32_t* csr_eid;    /* [E]    original edge id (row of x_e) at CSR position q      */
    const int32_t* csr_src;    /* [E]    fibre at CSR position q                              */
    const int32_t* csr_tgt;    /* [E]    class at CSR position q                              */
    const int32_t* tile_fibre; /* [ntiles+1] first fibre of every tile (whole fibres, <= PFS_TILE_EDGES edges) */
    int32_t ntiles;            /* tiles per graph (CSR layout; dense layout derives it)       */
    const int32_t* csc_colptr; /* [T+1]  first CSC position of each class                     */
    const int32_t* csc_q;      /* [E]    CSR position of the k-th class-sorted edge           */
} pfs_topology;

/* -------------------------------------------------------------------------------------------
 * misc
 * ---------------------------------------------------------------------------------------- */
int pfs_abi_version(void);
const char* pfs_last_error(void);
/* 1 if feature width F has compiled kernels in this build, else 0 */
int pfs_supports_fdim(int32_t F);
/* sizeof() of the argument structs, so a binding can check its own layout */
size_t pfs_sizeof_topology(void);
size_t pfs_sizeof_edge_args(void);
size_t pfs_sizeof_source_args(void);
size_t pfs_sizeof_target_args(void);
size_t pfs_sizeof_global_args(void);
size_t pfs_sizeof_head_args(void);

/* Launch accounting and optional per-kernel timing (used by bench.py; not part of the reference
 * surface).  pfs_launch_count: kernels launched by this library so far.  pfs_profile_enable(1)
 * starts recording a CUDA event after every kernel on its launch stream; pfs_profile_report waits
 * for them and writes one line "kernel launches total_ms" per kernel (returns bytes written). */
long long pfs_launch_count(void);
int pfs_profile_enable(int on);
int pfs_profile_report(char* buf, size_t buflen);

/* Workspace (bytes) sufficient for any forward/backward call on this topology. */
size_t pfs_workspace_bytes(const pfs_topology* topo);

/* -------------------------------------------------------------------------------------------
 * topology (replaces the implicit advanced-indexing / torch_scatter index handling of
 * reference src/gnn.py:98-100,135-144,187-190)
 * ---------------------------------------------------------------------------------------- */
/* Writes 1 to *flag_dev (device int32) when edge_index [2,E] int64 is exactly the canonical
 * dense order src = e / T, tgt = e % T with E = S*T, else 0. */
int pfs_detect_dense(const int64_t* edge_index, int64_t E, int32_t S, int32_t T, int32_t* flag_dev, void* stream);

/* Bytes of temporary device storage pfs_build_topology needs. */
size_t pfs_build_topology_temp_bytes(int64_t E, int32_t S, int32_t T);
/* Builds the int32 CSR/CSC arrays (stable counting sort, deterministic).  Outputs (device):
 * csr_rowptr [S+1], csr_eid [E], csr_src [E], csr_tgt [E], csc_colptr [T+1], csc_q [E],
 * tile_fibre [S+1] (capacity), *ntiles_dev and *max_degree_dev (device int32 each). */
int pfs_build_topology(const int64_t* edge_index, int64_t E, int32_t S, int32_t T,
                       int32_t* csr_rowptr, int32_t* csr_eid, int32_t* csr_src, int32_t* csr_tgt,
                       int32_t* csc_colptr, int32_t* csc_q, int32_t* tile_fibre,
                       int32_t* ntiles_dev, int32_t* max_degree_dev,
                       void* temp, size_t temp_bytes, void* stream);

/* -------------------------------------------------------------------------------------------
 * EdgeModel  (reference src/gnn.py:73-101: gather + cat + MLP(4F,4F,F) + BatchNorm1d applied twice)
 * ---------------------------------------------------------------------------------------- */
typedef struct pfs_edge_args {
    pfs_topology topo;
    /* inputs */
    const float *x_s, *x_t, *x_e, *u;            /* [G,S,F] [G,T,F] [G,E,F] [G,F] */
    const float *w1, *b1, *w2, *b2;              /* `0.weight` [4F,4F], `0.bias` [4F], `2.weight` [F,4F], `2.bias` [F] */
    const float *gamma, *beta;                   /* `norm.weight/bias` [F]; ignored when normed == 0 */
    float *running_mean, *running_var;           /* `norm.running_*` [F]; updated in place when training */
    int64_t* num_batches_tracked;                /* += 2 per graph when training (double norm)  */
    int32_t training, normed;
    float eps, momentum;
    /* forward outputs */
    float* x_e_out;                              /* [G,E,F] */
    float* bn_save;                              /* [G,4,F]: mean, biased var, combined scale, shift */
    /* backward inputs (x_e_out and bn_save as written by the forward) */
    const float* g_out;                          /* dL/dx_e_out [G,E,F] */
    /* backward outputs (overwritten) */
    float *g_x_s, *g_x_t, *g_x_e, *g_u;          /* [G,S,F] [G,T,F] [G,E,F] [G,F] */
    float *g_w1, *g_b1, *g_w2, *g_b2, *g_gamma, *g_beta;
    void* workspace; size_t workspace_bytes;
    void* stream;
} pfs_edge_args;
int pfs_edge_fwd(const pfs_edge_args* a);
int pfs_edge_bwd(const pfs_edge_args* a);

/* -------------------------------------------------------------------------------------------
 * SModel  (reference src/gnn.py:104-154: message MLP(2F,2F,2F), per-fibre mean/std/skew/kurtosis
 * via torch_scatter, node MLP(10F,10F,F), BatchNorm1d over the fibres)
 * ---------------------------------------------------------------------------------------- */
typedef struct pfs_source_args {
    pfs_topology topo;
    const float *x_s, *x_t, *x_e, *u;            /* x_e is the UPDATED edge embedding */
    const float *w1, *b1, *w2, *b2;              /* node_mlp_1: [2F,2F] [2F] [2F,2F] [2F] */
    const float *w3, *b3, *w4, *b4;              /* node_mlp_2: [10F,10F] [10F] [F,10F] [F] */
    const float *gamma, *beta;
    float *running_mean, *running_var;
    int64_t* num_batches_tracked;
    int32_t training, normed;
    float eps, momentum;
    float* x_s_out;                              /* [G,S,F] */
    float* moments;                              /* [G,S,5,2F] saved: mean, E[m^2], c2, c3, c4 */
    float* hidden;                               /* [G,S,10F] saved: lrelu of the node MLP hidden layer */
    float* y_pre;                                /* [G,S,F] saved: node MLP output before the norm */
    float* bn_save;                              /* [G,4,F] */
    const float* g_out;                          /* dL/dx_s_out [G,S,F] */
    float *g_x_s, *g_x_t, *g_x_e, *g_u;
    float *g_w1, *g_b1, *g_w2, *g_b2, *g_w3, *g_b3, *g_w4, *g_b4, *g_gamma, *g_beta;
    void* workspace; size_t workspace_bytes;
    void* stream;
} pfs_source_args;
int pfs_source_fwd(const pfs_source_args* a);
int pfs_source_bwd(const pfs_source_args* a);

/* -------------------------------------------------------------------------------------------
 * TModel  (reference src/gnn.py:157-192: message MLP(2F,2F,2F), scatter-sum over classes,
 * node MLP(4F,4F,F), BatchNorm1d over the classes)
 * ---------------------------------------------------------------------------------------- */
typedef struct pfs_target_args {
    pfs_topology topo;
    const float *x_s, *x_t, *x_e, *u;            /* x_s, x_e are the UPDATED embeddings */
    const float *w1, *b1, *w2, *b2;              /* node_mlp_1 */
    const float *w3, *b3, *w4, *b4;              /* node_mlp_2: [4F,4F] [4F] [F,4F] [F] */
    const float *gamma, *beta;
    float *running_mean, *running_var;
    int64_t* num_batches_tracked;
    int32_t training, normed;
    float eps, momentum;
    float* x_t_out;                              /* [G,T,F] */
    float* act_sum;                              /* [G,T,2F] saved: per-class sum of the hidden activations */
    float* y_pre;                                /* [G,T,F] saved */
    float* bn_save;                              /* [G,4,F] */
    const float* g_out;                          /* dL/dx_t_out [G,T,F] */
    float *g_x_s, *g_x_t, *g_x_e, *g_u;
    float *g_w1, *g_b1, *g_w2, *g_b2, *g_w3, *g_b3, *g_w4, *g_b4, *g_gamma, *g_beta;
    void* workspace; size_t workspace_bytes;
    void* stream;
} pfs_target_args;
int pfs_target_fwd(const pfs_target_args* a);
int pfs_target_bwd(const pfs_target_args* a);

/* -------------------------------------------------------------------------------------------
 * GlobalModel  (reference src/gnn.py:195-223: mean-pool, MLP(3F,3F,F), RMSNorm applied twice)
 * ---------------------------------------------------------------------------------------- */
typedef struct pfs_global_args {
    int32_t G, F, S, T;
    const float *x_s, *x_t, *u;                  /* [G,S,F] [G,T,F] [G,F] (updated node embeddings) */
    const float *w1, *b1, *w2, *b2;              /* [3F,3F] [3F] [F,3F] [F] */
    const float* rms_weight;                     /* `norm.weight` [F]; ignored when normed == 0 */
    int32_t normed;
    float rms_eps;                               /* torch.finfo(float32).eps for nn.RMSNorm(eps=None) */
    float* u_out;                                /* [G,F] */
    const float* g_out;                          /* dL/du_out [G,F] */
    float *g_x_s, *g_x_t, *g_u;                  /* [G,S,F] [G,T,F] [G,F] */
    float *g_w1, *g_b1, *g_w2, *g_b2, *g_rms_weight;
    void* workspace; size_t workspace_bytes;
    void* stream;
} pfs_global_args;
int pfs_global_fwd(const pfs_global_args* a);
int pfs_global_bwd(const pfs_global_args* a);

/* -------------------------------------------------------------------------------------------
 * Time head  (reference GNN.edge_prediction, src/gnn.py:307-312: MLP(F,F,1), round (identity,
 * src/gnn.py:321-325), softplus * scale) plus the integer times adopted in DESIGN.md:
 * visits = rint(time / hours[tgt]), time_int = visits * hours[tgt] (src/train.py:257).
 * ---------------------------------------------------------------------------------------- */
typedef struct pfs_head_args {
    pfs_topology topo;
    const float* x_e;                            /* [G,E,F] */
    const float *w1, *b1, *w2, *b2;              /* decoder_e: [F,F] [F] [1,F] [1] */
    float scale;
    const float* class_hours;                    /* [T] hours per visit, or NULL (no integer outputs) */
    const int64_t* edge_tgt;                     /* [E] class of every edge (edge_index[1]); NULL for the dense layout */
    float* time;                                 /* [G,E] */
    float* visits;                               /* [G,E] integer-valued, or NULL */
    float* time_int;                             /* [G,E] or NULL */
    const float* g_time;                         /* dL/dtime [G,E] */
    float* g_x_e;                                /* [G,E,F] */
    float *g_w1, *g_b1, *g_w2, *g_b2;
    void* workspace; size_t workspace_bytes;
    void* stream;
} pfs_head_args;
int pfs_time_head_fwd(const pfs_head_args* a);
int pfs_time_head_bwd(const pfs_head_args* a);

#ifdef __cplusplus
}
#endif
#endif /* PFS_B200_H */
