// node_ops.cuh -- node-level (per fibre / per class / per graph) kernels shared by the modules:
// small linear maps over node rows (the per-node tables of the first-layer split), their
// backward, BatchNorm finalisation / running-stat updates, and the fixed-order reductions of
// per-CTA partials.  All templated on compile-time widths; instantiated from api.cu.
#pragma once
#include "common.cuh"

namespace pfs {

// ------------------------------------------------------------------------------------------
// out[n][j] = bias[j] + addvec[g][j] + sum_k W[j][koff + k] * x[n][k]      (n = g * rows + r)
// thread-per-row, weights broadcast from shared memory.
// ------------------------------------------------------------------------------------------
template <int K, int J>
__global__ void __launch_bounds__(kThreads) k_node_linear(const float* __restrict__ x, int rows_per_graph, int G,
                                                           const float* __restrict__ W, int ldw, int koff,
                                                           const float* __restrict__ bias,
                                                           const float* __restrict__ addvec,
                                                           float* __restrict__ out) {
    __shared__ __align__(16) float Wt[K * J];
    __shared__ float bs[J];
    load_w_inmajor<K, J>(Wt, W, ldw, koff);
    for (int i = threadIdx.x; i < J; i += blockDim.x) bs[i] = bias ? __ldg(bias + i) : 0.f;
    __syncthreads();
    const long long N = (long long)rows_per_graph * G;
    for (long long n = (long long)blockIdx.x * kThreads + threadIdx.x; n < N; n += (long long)gridDim.x * kThreads) {
        float xr[K], y[J];
        load_row<K>(x + n * K, xr);
#pragma unroll
        for (int j = 0; j < J; ++j) y[j] = bs[j];
        if (addvec) {
            const int g = (int)(n / rows_per_graph);
            add_row<J>(addvec + (size_t)g * J, y);
        }
        dense_acc<K, J>(Wt, xr, y);
        store_row<J>(out + n * J, y);
    }
}

// ------------------------------------------------------------------------------------------
// Prologue of a module call in ONE launch: the per-node tables of the first-layer split (up to two: e.g. the
// EdgeModel's P_s = x_s . W1_s^T and P_t = x_t . W1_t^T + W1_u . u + b1) and the packing of the module's small MLP
// weights into the staging buffer of the constant bank.  The first `pack_blocks` CTAs pack, the others run the
// table rows thread-per-row with the weight slices broadcast from shared memory.
// ------------------------------------------------------------------------------------------
struct TableJob {
    const float* x;     // [G, rows, F]
    int rows;           // rows per graph
    int koff;           // first column of W contracted with x
    int ukoff;          // first column of W contracted with the graph's global row u[g] (added to every row), or -1
    const float* bias;  // [J] or null
    float* out;         // [G, rows, J]
};
struct PrepParams {
    TableJob job[2];
    int njobs, G;
    const float* W;     // [J, ldw] first-layer weight matrix
    int ldw;
    const float* u;     // [G, F]
    PackList pl;
    float* stage;
    int pack_blocks;
};
template <int F, int J>
__global__ void __launch_bounds__(kThreads) k_prep(const PrepParams p) {
    if ((int)blockIdx.x < p.pack_blocks) {
        for (int q = 0; q < p.pl.n; ++q) {
            const PackItem& it = p.pl.it[q];
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < it.J * it.K; i += p.pack_blocks * blockDim.x) {
                const int j = i / it.K, k = i - j * it.K;
                const float v = __ldg(it.W + (size_t)j * it.ld + it.koff + k);
                p.stage[it.dst_off + (it.transpose ? k * it.J + j : j * it.K + k)] = v;
            }
        }
        return;
    }
    __shared__ __align__(16) float Wt[2][F * J];
    __shared__ __align__(16) float Wu[F * J];
    __shared__ float bs[2][J];
    const int nb = gridDim.x - p.pack_blocks, b = blockIdx.x - p.pack_blocks;
    bool any_u = false;
    for (int q = 0; q < p.njobs; ++q) {
        load_w_inmajor<F, J>(Wt[q], p.W, p.ldw, p.job[q].koff);
        for (int i = threadIdx.x; i < J; i += blockDim.x) bs[q][i] = p.job[q].bias ? __ldg(p.job[q].bias + i) : 0.f;
        if (p.job[q].ukoff >= 0) {
            load_w_inmajor<F, J>(Wu, p.W, p.ldw, p.job[q].ukoff);
            any_u = true;
        }
    }
    (void)any_u;
    __syncthreads();
    for (int q = 0; q < p.njobs; ++q) {
        const TableJob& jb = p.job[q];
        const long long N = (long long)jb.rows * p.G;
        for (long long n = (long long)b * kThreads + threadIdx.x; n < N; n += (long long)nb * kThreads) {
            float xr[F], y[J];
            load_row<F>(jb.x + n * F, xr);
#pragma unroll
            for (int j = 0; j < J; ++j) y[j] = bs[q][j];
            if (jb.ukoff >= 0) {
                float ur[F];
                load_row<F>(p.u + (n / jb.rows) * F, ur);
                dense_acc<F, J>(Wu, ur, y);
            }
            dense_acc<F, J>(Wt[q], xr, y);
            store_row<J>(jb.out + n * J, y);
        }
    }
}

// dx[n][k] (+)= sum_j W[j][koff + k] * d[n][j]     (backward of the map above w.r.t. x)
template <int K, int J, bool ACCUM>
__global__ void __launch_bounds__(kThreads) k_node_linear_bwd(const float* __restrict__ d, long long N,
                                                               const float* __restrict__ W, int ldw, int koff,
                                                               float* __restrict__ dx) {
    __shared__ __align__(16) float Wo[J * K];  // Wo[j*K + k]: "input" index j, "output" index k
    load_w_outmajor<K, J>(Wo, W, ldw, koff);
    __syncthreads();
    for (long long n = (long long)blockIdx.x * kThreads + threadIdx.x; n < N; n += (long long)gridDim.x * kThreads) {
        float dr[J], y[K];
        load_row<J>(d + n * J, dr);
        if (ACCUM) load_row<K>(dx + n * K, y);
        else {
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] = 0.f;
        }
        dense_acc<J, K>(Wo, dr, y);
        store_row<K>(dx + n * K, y);
    }
}

// ------------------------------------------------------------------------------------------
// Backward of a node table  P = X . W[:, koff : koff + K]^T  given D = dL/dP, one pass over D:
//   dW[j][k] partial = sum over a CTA's rows of D[n][j] * X[n][k]; bias partial = column sums of D;
//   DX: dx[n][k] = sum_j W[j][koff + k] * D[n][j]   (what k_node_linear_bwd computes in a pass of its own).
// Persistent CTAs over row tiles of kTile rows; partial p of CTA c at out[c * pstride + ...].
// Layout inside a CTA partial: [J*K] weights (row-major j, k) then [J] column sums.
// D tiles are staged with 16-byte loads into 16-byte aligned rows (J % 4 == 0), X tiles with 8-byte loads.
// ------------------------------------------------------------------------------------------
template <int J, int K, int TJ, int TK, bool DX>
__global__ void __launch_bounds__(kThreads) k_outer_rows(const float* __restrict__ D, const float* __restrict__ X,
                                                          long long N, float* __restrict__ partial, int pstride,
                                                          const float* __restrict__ W, int ldw, int koff,
                                                          float* __restrict__ dx) {
    static_assert(J % 4 == 0 && K % 2 == 0, "16-byte rows of D, 8-byte rows of X");
    constexpr int LDD = J + 4, LDX = K + 2;
    using Acc = OuterAcc<J, K, TJ, TK>;
    constexpr int AD = (TJ % 4 == 0) ? 4 : (TJ % 2 == 0) ? 2 : 1;
    constexpr int AX = (LDX % 4 == 0 && TK % 4 == 0) ? 4 : (TK % 2 == 0) ? 2 : 1;
    extern __shared__ __align__(16) float sm[];
    float* Ds = sm;
    float* Xs = sm + kTile * LDD;
    float* Wo = Xs + kTile * LDX;          // DX: Wo[j * K + k] = W[j][koff + k]
    if constexpr (DX) load_w_outmajor<K, J>(Wo, W, ldw, koff);
    Acc acc;
    acc.init();
    float colsum = 0.f;  // thread j < J accumulates column j
    const long long ntile = (N + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long n0 = tile * kTile;
        const int rows = (int)min((long long)kTile, N - n0);
        {
            constexpr int J4 = J / 4, K2 = K / 2;
            const float4* src = reinterpret_cast<const float4*>(D + n0 * J);
            for (int i = threadIdx.x; i < rows * J4; i += kThreads) {
                const int r = i / J4, c = i - r * J4;
                *reinterpret_cast<float4*>(Ds + r * LDD + 4 * c) = __ldg(src + i);
            }
            const float2* srx = reinterpret_cast<const float2*>(X + n0 * K);
            for (int i = threadIdx.x; i < rows * K2; i += kThreads) {
                const int r = i / K2, c = i - r * K2;
                *reinterpret_cast<float2*>(Xs + r * LDX + 2 * c) = __ldg(srx + i);
            }
        }
        __syncthreads();
        if constexpr (DX) {
            if ((int)threadIdx.x < rows) {
                float dr[J], y[K];
                lds_row<J>(Ds + threadIdx.x * LDD, dr);
#pragma unroll
                for (int k = 0; k < K; ++k) y[k] = 0.f;
                dense_acc<J, K>(Wo, dr, y);
                store_row<K>(dx + (n0 + threadIdx.x) * K, y);
            }
        }
        acc.template accumulate<AD, AX>(Ds, LDD, Xs, LDX, rows);
        if (threadIdx.x < J) {
            float s0 = 0.f, s1 = 0.f;
            int r = 0;
            for (; r + 1 < rows; r += 2) {
                s0 += Ds[r * LDD + threadIdx.x];
                s1 += Ds[(r + 1) * LDD + threadIdx.x];
            }
            if (r < rows) s0 += Ds[r * LDD + threadIdx.x];
            colsum += s0 + s1;
        }
        __syncthreads();
    }
    float* out = partial + (size_t)blockIdx.x * pstride;
    acc.flush(sm, out, K, 0);
    if (threadIdx.x < J) out[J * K + threadIdx.x] = colsum;
}
template <int J, int K, int TJ, int TK>
constexpr size_t outer_rows_smem() {
    constexpr int LDD = J + 4, LDX = K + 2;
    constexpr int a = kTile * (LDD + LDX) + J * K;
    constexpr int b = OuterAcc<J, K, TJ, TK>::kScratchFloats;
    return sizeof(float) * (a > b ? a : b);
}

// out[i] = sum_c partial[c * pstride + poff + i] for i < n  (fixed order over the CTAs);
// 2-D destination: out[(i / cols) * ldo + coff + (i % cols)].
__global__ void k_reduce_partials(const float* __restrict__ partial, int ncta, int pstride, int poff, int n, int cols,
                                  float* __restrict__ out, int ldo, int coff) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int c = 0; c < ncta; ++c) s += partial[(size_t)c * pstride + poff + i];
    out[(i / cols) * ldo + coff + (i % cols)] = s;
}

// several reductions in one launch (fewer latency-bound launches): every segment names its own partial buffer, so
// all the fixed-order final sums of a module call (edge-kernel partials, node-kernel partials, table-gradient
// partials) go out together at the end of the call
struct ReduceSeg {
    const float* partial;
    int ncta, pstride;
    int poff, n, cols, ldo, coff;
    float* out;
};
constexpr int kMaxReduceSegs = 12;
struct ReduceList {
    ReduceSeg seg[kMaxReduceSegs];
    int nseg, total;
};
// block = kReduceSlices slices x 32 outputs: slice s adds the partials of CTAs s, s + slices, ... (one short
// dependent chain each instead of one chain over all CTAs), the slices are then added in a fixed order
constexpr int kReduceSlices = 8;
__global__ void __launch_bounds__(32 * kReduceSlices) k_reduce_multi(const ReduceList rl) {
    __shared__ float red[kReduceSlices][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    int i = blockIdx.x * 32 + lane;
    const bool live = i < rl.total;
    int q = 0;
    while (live && q < rl.nseg - 1 && i >= rl.seg[q].n) {
        i -= rl.seg[q].n;
        ++q;
    }
    float s0 = 0.f, s1 = 0.f;        // fixed association: two interleaved chains per slice
    if (live) {
        const float* p = rl.seg[q].partial + rl.seg[q].poff + i;
        const int ncta = rl.seg[q].ncta;
        const size_t pstride = (size_t)rl.seg[q].pstride;
        int c = slice;
        for (; c + kReduceSlices < ncta; c += 2 * kReduceSlices) {
            s0 += p[(size_t)c * pstride];
            s1 += p[(size_t)(c + kReduceSlices) * pstride];
        }
        if (c < ncta) s0 += p[(size_t)c * pstride];
    }
    red[slice][lane] = s0 + s1;
    __syncthreads();
    if (slice == 0 && live) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kReduceSlices; ++k) s += red[k][lane];
        const ReduceSeg& sg = rl.seg[q];
        sg.out[(i / sg.cols) * sg.ldo + sg.coff + (i % sg.cols)] = s;
    }
}

// colsum over rows and graphs: out[j] = sum_n x[n][j] (single CTA per 32 columns; fixed order)
__global__ void k_colsum_all(const float* __restrict__ x, long long N, int ld, int off, int J,
                             float* __restrict__ out) {
    // blockDim = (32 columns, 8 row lanes); column j of the result is x[:, off + j] of a [N, ld] matrix
    __shared__ float red[8][33];
    const int j = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (j < J)
        for (long long n = threadIdx.y; n < N; n += 8) s += x[n * ld + off + j];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && j < J) {
        float t = 0.f;
        for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
        out[j] = t;
    }
}

// the same for two column blocks in one launch (blockIdx.y picks the block): the gamma / beta gradient pairs
__global__ void k_colsum_pair(const float* __restrict__ x, long long N, int ld, int off0, float* __restrict__ out0, int off1,
                              float* __restrict__ out1, int J) {
    __shared__ float red[8][33];
    const int off = blockIdx.y ? off1 : off0;
    float* out = blockIdx.y ? out1 : out0;
    const int j = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (j < J)
        for (long long n = threadIdx.y; n < N; n += 8) s += x[n * ld + off + j];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && j < J) {
        float t = 0.f;
        for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
        out[j] = t;
    }
}

// per-graph column sums: out[g][j] = sum_r x[g][r][j]
__global__ void k_colsum_graph(const float* __restrict__ x, int rows, int J, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int g = blockIdx.y;
    const int j = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (j < J)
        for (int r = threadIdx.y; r < rows; r += 8) s += x[((size_t)g * rows + r) * J + j];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && j < J) {
        float t = 0.f;
        for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
        out[(size_t)g * J + j] = t;
    }
}

// out[j][koff + k] = sum_g a[g][j] * b[g][k]      (the `u` columns of a first-layer weight gradient)
// block = 8 slices x 32 outputs: slice s takes graphs s, s + 8, ... with four loads in flight, the slices are
// added in a fixed order (one thread looping over all graphs was a 256-deep chain of dependent L2 round trips)
__global__ void __launch_bounds__(256) k_outer_graphs(const float* __restrict__ a, const float* __restrict__ b, int G,
                                                      int J, int K, float* __restrict__ out, int ldo, int koff) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const bool live = i < J * K;
    const int j = live ? i / K : 0, k = live ? i - j * K : 0;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int g = slice;
    for (; g + 24 < G; g += 32) {
        s0 = fmaf(a[(size_t)g * J + j], b[(size_t)g * K + k], s0);
        s1 = fmaf(a[(size_t)(g + 8) * J + j], b[(size_t)(g + 8) * K + k], s1);
        s2 = fmaf(a[(size_t)(g + 16) * J + j], b[(size_t)(g + 16) * K + k], s2);
        s3 = fmaf(a[(size_t)(g + 24) * J + j], b[(size_t)(g + 24) * K + k], s3);
    }
    for (; g < G; g += 8) s0 = fmaf(a[(size_t)g * J + j], b[(size_t)g * K + k], s0);
    red[slice][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (slice == 0 && live) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red[q][lane];
        out[j * ldo + koff + k] = s;
    }
}

// ------------------------------------------------------------------------------------------
// class-side reduction of per-tile partial sums: out[g][i][j] = sum_lt part[g][lt][i][j]
// ------------------------------------------------------------------------------------------
__global__ void k_class_reduce(const float* __restrict__ part, int ntiles, int TJ, float* __restrict__ out) {
    const int g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TJ) return;
    const float* p = part + (size_t)g * ntiles * TJ + i;
    float s = 0.f;
    for (int t = 0; t < ntiles; ++t) s += p[(size_t)t * TJ];
    out[(size_t)g * TJ + i] = s;
}

// general layout: out[g][c][j] = sum over the class's edges (class-sorted order) of rows[g][q][j]
// grid (T, nchunk, G), 128 threads: a block sums one chunk of one class's edge list; thread = (row lane,
// feature), row lanes summed in a fixed order through shared memory.  nchunk > 1 writes partials
// [g][chunk][c][J] for k_csc_segment_final (a few hundred long class segments must fill the device).
constexpr int kCscThreads = 128;
__global__ void __launch_bounds__(kCscThreads) k_csc_segment_sum(const float* __restrict__ rows, const int* __restrict__ colptr,
                                                                 const int* __restrict__ cscq, int E, int T, int J, int nchunk,
                                                                 float* __restrict__ out, float* __restrict__ partial) {
    __shared__ float red[kCscThreads];
    const int c = blockIdx.x, chunk = blockIdx.y, g = blockIdx.z;
    const int a = colptr[c], len = colptr[c + 1] - a;
    const int i0 = a + (int)((long long)len * chunk / nchunk), i1 = a + (int)((long long)len * (chunk + 1) / nchunk);
    const int lanes = kCscThreads / J;                       // J <= 128
    const int lr = threadIdx.x / J, j = threadIdx.x - lr * J;
    const float* base = rows + (size_t)g * E * J + j;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (lr < lanes) {
        int k = i0 + lr;
        for (; k + 3 * lanes < i1; k += 4 * lanes) {         // four gathers in flight per thread
            const int q0 = cscq[k], q1 = cscq[k + lanes], q2 = cscq[k + 2 * lanes], q3 = cscq[k + 3 * lanes];
            s0 += base[(size_t)q0 * J]; s1 += base[(size_t)q1 * J];
            s2 += base[(size_t)q2 * J]; s3 += base[(size_t)q3 * J];
        }
        for (; k < i1; k += lanes) s0 += base[(size_t)cscq[k] * J];
    }
    red[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (lr == 0) {
        float s = 0.f;
        for (int r = 0; r < lanes; ++r) s += red[r * J + j];
        if (nchunk > 1) partial[(((size_t)g * nchunk + chunk) * T + c) * J + j] = s;
        else out[((size_t)g * T + c) * J + j] = s;
    }
}
__global__ void k_csc_segment_final(const float* __restrict__ partial, int nchunk, int TJ, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (i >= TJ) return;
    float s = 0.f;
    for (int q = 0; q < nchunk; ++q) s += partial[((size_t)g * nchunk + q) * TJ + i];
    out[(size_t)g * TJ + i] = s;
}

// sums of the per-tile BatchNorm-backward partials of a graph: out[g][0..2F) in double, fixed order
// (one CTA per graph, threads stride over the tiles: a single large graph has tens of thousands of tiles)
__global__ void __launch_bounds__(256) k_tile_partial_sums(const float* __restrict__ partial, int ntiles, int F2,
                                                           double* __restrict__ out) {
    __shared__ double red[256];
    const int g = blockIdx.x;
    const int lanes = 256 / F2;
    const int lr = threadIdx.x / F2, f = threadIdx.x - lr * F2;
    double s = 0.0;
    if (lr < lanes)
        for (int t = lr; t < ntiles; t += lanes) s += (double)partial[((size_t)g * ntiles + t) * F2 + f];
    red[threadIdx.x] = s;
    __syncthreads();
    if (lr == 0) {
        double tot = 0.0;
        for (int r = 0; r < lanes; ++r) tot += red[r * F2 + f];
        out[(size_t)g * F2 + f] = tot;
    }
}

// ------------------------------------------------------------------------------------------
// BatchNorm: combine tile partials (Chan, fp64, fixed order), closed-form coefficients
//   save[g][0][f] = mean, save[g][1][f] = biased variance, save[g][2][f] = scale, save[g][3][f] = shift
//   twice == 1: the edge model's double application (reference src/gnn.py:82,101):
//       scale = gamma^2 r1 r2, shift = beta - scale * mean, r2 = rsqrt(gamma^2 var r1^2 + eps)
// ------------------------------------------------------------------------------------------
__global__ void k_bn_finalize(const float* __restrict__ partial, int ntiles, int pstride, int F, int G,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int twice,
                              float* __restrict__ save) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * F) return;
    const int g = i / F, f = i - g * F;
    const float* p = partial + (size_t)g * ntiles * pstride;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        const double nb = p[(size_t)t * pstride + 2 * F];
        if (nb <= 0.0) continue;
        const double mb = p[(size_t)t * pstride + f], sb = p[(size_t)t * pstride + F + f];
        const double tot = n + nb, delta = mb - mean;
        mean += delta * (nb / tot);
        m2 += sb + delta * delta * (n * nb / tot);
        n = tot;
    }
    const double var = n > 0 ? m2 / n : 0.0;
    const double gm = gamma[f], bt = beta[f];
    const double r1 = 1.0 / sqrt(var + (double)eps);
    double scale, shift;
    if (twice) {
        const double var2 = gm * gm * var * r1 * r1;
        const double r2 = 1.0 / sqrt(var2 + (double)eps);
        scale = gm * gm * r1 * r2;
    } else {
        scale = gm * r1;
    }
    shift = bt - scale * mean;
    float* s = save + (size_t)g * 4 * F;
    s[f] = (float)mean;
    s[F + f] = (float)var;
    s[2 * F + f] = (float)scale;
    s[3 * F + f] = (float)shift;
}

// same, from the per-CTA graph-segment records of the chunked edge kernels:
// partial[cta][rec] = {mean[F], M2[F], count, graph}; CTA c owns tiles [c*total/ncta, (c+1)*total/ncta)
__global__ void k_bn_finalize_rec(const float* __restrict__ partial, int ncta, int nrec, int tiles_per_graph, int F,
                                  int G, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                  int twice, float* __restrict__ save) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * F) return;
    const int g = i / F, f = i - g * F;
    const long long total = (long long)tiles_per_graph * G;
    long long c_lo = (long long)g * tiles_per_graph * ncta / total - 1;
    long long c_hi = ((long long)(g + 1) * tiles_per_graph * ncta) / total + 1;
    if (c_lo < 0) c_lo = 0;
    if (c_hi > ncta - 1) c_hi = ncta - 1;
    const int stride = 2 * F + 2;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (long long c = c_lo; c <= c_hi; ++c)
        for (int r = 0; r < nrec; ++r) {
            const float* p = partial + ((size_t)c * nrec + r) * stride;
            const double nb = p[2 * F];
            if (nb <= 0.0 || (int)p[2 * F + 1] != g) continue;
            const double mb = p[f], sb = p[F + f];
            const double tot = n + nb, delta = mb - mean;
            mean += delta * (nb / tot);
            m2 += sb + delta * delta * (n * nb / tot);
            n = tot;
        }
    const double var = n > 0 ? m2 / n : 0.0;
    const double gm = gamma[f], bt = beta[f];
    const double r1 = 1.0 / sqrt(var + (double)eps);
    double scale;
    if (twice) {
        const double var2 = gm * gm * var * r1 * r1;
        scale = gm * gm * r1 / sqrt(var2 + (double)eps);
    } else {
        scale = gm * r1;
    }
    float* s = save + (size_t)g * 4 * F;
    s[f] = (float)mean;
    s[F + f] = (float)var;
    s[2 * F + f] = (float)scale;
    s[3 * F + f] = (float)(bt - scale * mean);
}

// finalisation (either partial format) AND the running-buffer update in one single-CTA launch: the per-graph
// statistics are a few thousand values, two dependent latency-bound launches cost more than the work
struct BnFinalizeAll {
    const float* partial;
    int rec_ncta, rec_nrec;      // rec_ncta > 0: per-CTA graph-segment records, else per-tile partials
    int ntiles, pstride, F, G, twice;
    const float *gamma, *beta;
    float eps, momentum;
    long long n_rows;
    float* save;
    float *rm, *rv;
    long long* nbt;
};
// ticket of the last-CTA pattern below (single-stream contract of the library: one such kernel at a time per device)
__device__ unsigned int g_bn_ticket = 0;

// Chan-merges up to kBatch records whose loads were issued together (the finalisation is latency-bound: a chain of
// dependent L2 round trips per record made the single-CTA version take 25 - 50 us)
__global__ void __launch_bounds__(256) k_bn_finalize_all(const BnFinalizeAll a) {
    constexpr int kBatch = 4;
    const int F = a.F, G = a.G;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G * F) {
        const int g = i / F, f = i - g * F;
        double n = 0.0, mean = 0.0, m2 = 0.0;
        const float* base;
        long long stride;
        int count;
        if (a.rec_ncta > 0) {
            const long long total = (long long)a.ntiles * G;
            long long c_lo = (long long)g * a.ntiles * a.rec_ncta / total - 1;
            long long c_hi = ((long long)(g + 1) * a.ntiles * a.rec_ncta) / total + 1;
            if (c_lo < 0) c_lo = 0;
            if (c_hi > a.rec_ncta - 1) c_hi = a.rec_ncta - 1;
            stride = 2 * F + 2;
            base = a.partial + (size_t)c_lo * a.rec_nrec * stride;      // records of CTAs c_lo .. c_hi are contiguous
            count = (int)(c_hi - c_lo + 1) * a.rec_nrec;
        } else {
            stride = a.pstride;
            base = a.partial + (size_t)g * a.ntiles * a.pstride;
            count = a.ntiles;
        }
        for (int t0 = 0; t0 < count; t0 += kBatch) {
            float nb[kBatch], gid[kBatch], mb[kBatch], sb[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const bool in = t0 + k < count;
                const float* p = base + (size_t)(in ? t0 + k : t0) * stride;
                nb[k] = in ? p[2 * F] : 0.f;
                gid[k] = a.rec_ncta > 0 ? p[2 * F + 1] : (float)g;
                mb[k] = p[f];
                sb[k] = p[F + f];
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                if (!(nb[k] > 0.f) || (int)gid[k] != g) continue;
                const double cnt = nb[k], tot = n + cnt, delta = (double)mb[k] - mean;
                mean += delta * (cnt / tot);
                m2 += (double)sb[k] + delta * delta * (n * cnt / tot);
                n = tot;
            }
        }
        const double var = n > 0 ? m2 / n : 0.0;
        const double gm = a.gamma[f], bt = a.beta[f];
        const double r1 = 1.0 / sqrt(var + (double)a.eps);
        double scale;
        if (a.twice) {
            const double var2 = gm * gm * var * r1 * r1;
            scale = gm * gm * r1 / sqrt(var2 + (double)a.eps);
        } else {
            scale = gm * r1;
        }
        float* s = a.save + (size_t)g * 4 * F;
        s[f] = (float)mean;
        s[F + f] = (float)var;
        s[2 * F + f] = (float)scale;
        s[3 * F + f] = (float)(bt - scale * mean);
    }
    if (!a.rm || !a.rv) return;
    // running buffers: by the CTA that finishes last, after every CTA's statistics are visible
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&g_bn_ticket, 1u);
        is_last = t == gridDim.x - 1;
        if (is_last) g_bn_ticket = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31;
    const double mom = a.momentum, keep = 1.0 - mom, unb = (double)a.n_rows / (double)(a.n_rows - 1);
    const double step = a.twice ? keep * keep : keep;
    for (int f = threadIdx.x >> 5; f < F; f += blockDim.x >> 5) {
        double am = 0.0, av = 0.0;
        for (int g = lane; g < G; g += 32) {
            const double mean = __ldcg(a.save + (size_t)g * 4 * F + f), var = __ldcg(a.save + (size_t)g * 4 * F + F + f);
            double cm, cv;
            if (a.twice) {
                const double gm = a.gamma[f];
                const double var2 = gm * gm * var / (var + (double)a.eps);
                cm = keep * mom * mean + mom * (double)a.beta[f];
                cv = keep * mom * var * unb + mom * var2 * unb;
            } else {
                cm = mom * mean;
                cv = mom * var * unb;
            }
            const double w = pow(step, (double)(G - 1 - g));
            am += w * cm;
            av += w * cv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            am += __shfl_xor_sync(0xffffffffu, am, o);
            av += __shfl_xor_sync(0xffffffffu, av, o);
        }
        if (lane == 0) {
            const double d = pow(step, (double)G);
            a.rm[f] = (float)(d * (double)a.rm[f] + am);
            a.rv[f] = (float)(d * (double)a.rv[f] + av);
            if (f == 0 && a.nbt) *a.nbt += (long long)G * (a.twice ? 2 : 1);
        }
    }
}

// eval mode: coefficients from the running buffers (same for every graph)
__global__ void k_bn_eval_coeffs(const float* __restrict__ rm, const float* __restrict__ rv, int F, int G,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int twice,
                                 float* __restrict__ save) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * F) return;
    const int g = i / F, f = i - g * F;
    const double a = (double)gamma[f] / sqrt((double)rv[f] + (double)eps);
    const double m = rm[f], bt = beta[f];
    double scale, shift;
    if (twice) {  // y2 = a (a (z - m) + bt - m) + bt
        scale = a * a;
        shift = a * (bt - m - a * m) + bt;
    } else {
        scale = a;
        shift = bt - a * m;
    }
    float* s = save + (size_t)g * 4 * F;
    s[f] = (float)m;
    s[F + f] = rv[f];
    s[2 * F + f] = (float)scale;
    s[3 * F + f] = (float)shift;
}

// running buffers, graph after graph (what G sequential reference forwards would leave behind):
// r_G = (1-m)^(kG) r_0 + sum_g (weights) -- a fixed-order weighted sum, one warp per feature.
__global__ void k_bn_running(const float* __restrict__ save, int F, int G, long long n_rows,
                             const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                             float momentum, int twice, float* __restrict__ rm, float* __restrict__ rv,
                             long long* __restrict__ nbt) {
    const int f = blockIdx.x, lane = threadIdx.x;   // blockDim.x == 32
    const double mom = momentum, keep = 1.0 - mom, unb = (double)n_rows / (double)(n_rows - 1);
    const double step = twice ? keep * keep : keep;   // decay per graph
    double am = 0.0, av = 0.0;
    for (int g = lane; g < G; g += 32) {
        const double mean = save[(size_t)g * 4 * F + f], var = save[(size_t)g * 4 * F + F + f];
        double cm, cv;   // contribution of graph g right after its own update(s)
        if (twice) {     // second application sees mean beta and variance gamma^2 var / (var + eps)
            const double gm = gamma[f];
            const double var2 = gm * gm * var / (var + (double)eps);
            cm = keep * mom * mean + mom * (double)beta[f];
            cv = keep * mom * var * unb + mom * var2 * unb;
        } else {
            cm = mom * mean;
            cv = mom * var * unb;
        }
        const double w = pow(step, (double)(G - 1 - g));
        am += w * cm;
        av += w * cv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        am += __shfl_xor_sync(0xffffffffu, am, o);
        av += __shfl_xor_sync(0xffffffffu, av, o);
    }
    if (lane == 0) {
        const double d = pow(step, (double)G);
        rm[f] = (float)(d * (double)rm[f] + am);
        rv[f] = (float)(d * (double)rv[f] + av);
        if (f == 0 && nbt) *nbt += (long long)G * (twice ? 2 : 1);
    }
}

// out[g][r][f] = in[g][r][f] * scale[g][f] + shift[g][f]   (in place allowed); rows_f = rows * F (even)
__global__ void k_affine_rows(const float* __restrict__ in, const float* __restrict__ save, int F, long long rows_f,
                              int G, float* __restrict__ out) {
    // grid.y = graph; each thread handles float2 pairs of one graph
    const int g = blockIdx.y;
    const float* s = save + (size_t)g * 4 * F;
    const float2* src = reinterpret_cast<const float2*>(in + (size_t)g * rows_f);
    float2* dst = reinterpret_cast<float2*>(out + (size_t)g * rows_f);
    const long long n2 = rows_f / 2;
    const int half = F / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const int f = 2 * (int)(i % half);
        float2 v = src[i];
        v.x = fmaf(v.x, s[2 * F + f], s[3 * F + f]);
        v.y = fmaf(v.y, s[2 * F + f + 1], s[3 * F + f + 1]);
        dst[i] = v;
    }
}

// BatchNorm backward statistics over node rows: per graph sum_r g and sum_r g * xhat,
// xhat = (y - mean) * rsqrt(var + eps).  One CTA per graph (rows <= a few thousand per graph).
// out[g][0][f] = sum g, out[g][1][f] = sum g * xhat
// grid (nchunk, G): chunk c of graph g covers rows [rows*c/nchunk, rows*(c+1)/nchunk); nchunk == 1 writes
// out[g][2F] directly, otherwise partial[g][chunk][2F] for k_bn_bwd_stats_final
__global__ void k_bn_bwd_stats_rows(const float* __restrict__ gout, const float* __restrict__ y,
                                    const float* __restrict__ save, int rows, int F, float eps, int nchunk,
                                    float* __restrict__ out) {
    // thread = (row lane, feature): coalesced single pass over the rows, row lanes summed in order
    extern __shared__ float red[];  // [2][blockDim.x]
    const int chunk = blockIdx.x, g = blockIdx.y;
    const int lanes = blockDim.x / F;
    const int lr = threadIdx.x / F, f = threadIdx.x - lr * F;
    const float* s = save + (size_t)g * 4 * F;
    const int r_begin = (int)((long long)rows * chunk / nchunk), r_end = (int)((long long)rows * (chunk + 1) / nchunk);
    float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
    if (lr < lanes) {
        const float mean = s[f], r = rsqrtf(s[F + f] + eps);
        int rr = r_begin + lr;
        // four rows (eight loads) in flight per thread: the pass is a chain of memory round trips, not bandwidth
        for (; rr + 3 * lanes < r_end; rr += 4 * lanes) {
            const size_t i0 = ((size_t)g * rows + rr) * F + f, st = (size_t)lanes * F;
            const float g0 = gout[i0], g1 = gout[i0 + st], g2 = gout[i0 + 2 * st], g3 = gout[i0 + 3 * st];
            const float y0 = y[i0], y1 = y[i0 + st], y2 = y[i0 + 2 * st], y3 = y[i0 + 3 * st];
            a0 += g0; a1 += g1;
            b0 += g0 * ((y0 - mean) * r);
            b1 += g1 * ((y1 - mean) * r);
            a0 += g2; a1 += g3;
            b0 += g2 * ((y2 - mean) * r);
            b1 += g3 * ((y3 - mean) * r);
        }
        for (; rr + lanes < r_end; rr += 2 * lanes) {
            const size_t i0 = ((size_t)g * rows + rr) * F + f, i1 = i0 + (size_t)lanes * F;
            const float g0 = gout[i0], g1 = gout[i1], y0 = y[i0], y1 = y[i1];
            a0 += g0; a1 += g1;
            b0 += g0 * ((y0 - mean) * r);
            b1 += g1 * ((y1 - mean) * r);
        }
        if (rr < r_end) {
            const size_t i0 = ((size_t)g * rows + rr) * F + f;
            const float g0 = gout[i0];
            a0 += g0;
            b0 += g0 * ((y[i0] - mean) * r);
        }
    }
    red[threadIdx.x] = a0 + a1;
    red[blockDim.x + threadIdx.x] = b0 + b1;
    __syncthreads();
    if (threadIdx.x < 2 * F) {
        const int which = threadIdx.x / F, ff = threadIdx.x - which * F;
        float t = 0.f;
        for (int i = 0; i < lanes; ++i) t += red[which * blockDim.x + i * F + ff];
        out[((size_t)g * nchunk + chunk) * 2 * F + threadIdx.x] = t;
    }
}
__global__ void k_bn_bwd_stats_final(const float* __restrict__ partial, int nchunk, int F2, int G, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * F2) return;
    const int g = i / F2, f = i - g * F2;
    float s = 0.f;
    for (int c = 0; c < nchunk; ++c) s += partial[((size_t)g * nchunk + c) * F2 + f];
    out[i] = s;
}

}  // namespace pfs
