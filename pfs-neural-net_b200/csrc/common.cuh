// common.cuh -- device-side building blocks shared by every kernel of libpfs_b200.so (sm_100a).
//
// Thread mapping used throughout: a CTA of kThreads threads owns one tile of
// at most kTile fibre-sorted edges made of WHOLE fibres.  Per-edge work (the small MLPs) is
// thread-per-edge with the whole row in registers and the weights broadcast from shared memory;
// per-fibre work (moments, fibre sums) and per-class work (class sums) switch to
// thread-per-(segment, feature) over a shared-memory staging of the tile, so every reduction has
// a fixed order: no atomics anywhere, results are bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pfs_b200.h"

namespace pfs {

constexpr int kTile = PFS_TILE_EDGES;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr float kSlope = 0.1f;      // LeakyReLU(0.1) of every MLP (reference src/gnn.py:69)
constexpr float kSlopeVar = 0.01f;  // F.leaky_relu default on the variance (reference src/gnn.py:141)
constexpr float kStdEps = 1e-6f;    // reference src/gnn.py:142,149
constexpr int kNumSM = 148;

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : kSlope * x; }
__device__ __forceinline__ float dlrelu(float x) { return x > 0.f ? 1.f : kSlope; }

// ------------------------------------------------------------------------------------------
// topology
// ------------------------------------------------------------------------------------------
struct Topo {
    int layout, G, F, S, T, E;
    const int *rowptr, *eid, *csrc, *ctgt, *tile_fibre;
    int ntiles;  // tiles per graph
    const int *colptr, *cscq;
    int fpt;     // fibres per tile (dense layout)
};

struct Tile {
    int g, lt, fibre0, nfib, q0, ne;
};

__device__ __forceinline__ Tile get_tile(const Topo& tp, int tile) {
    Tile t;
    t.g = tile / tp.ntiles;
    t.lt = tile - t.g * tp.ntiles;
    if (tp.layout == PFS_LAYOUT_DENSE) {
        t.fibre0 = t.lt * tp.fpt;
        t.nfib = min(tp.fpt, tp.S - t.fibre0);
        t.q0 = t.fibre0 * tp.T;
        t.ne = t.nfib * tp.T;
    } else {
        t.fibre0 = tp.tile_fibre[t.lt];
        const int f1 = tp.tile_fibre[t.lt + 1];
        t.nfib = f1 - t.fibre0;
        t.q0 = tp.rowptr[t.fibre0];
        t.ne = tp.rowptr[f1] - t.q0;
    }
    return t;
}

struct EdgeRef {
    int e, src, tgt;  // row of x_e, fibre, class (all per graph)
};

__device__ __forceinline__ EdgeRef get_edge(const Topo& tp, const Tile& t, int tid) {
    EdgeRef r;
    const int q = t.q0 + tid;
    if (tp.layout == PFS_LAYOUT_DENSE) {
        const int lf = tid / tp.T;
        r.src = t.fibre0 + lf;
        r.tgt = tid - lf * tp.T;
        r.e = q;
    } else {
        r.src = tp.csrc[q];
        r.tgt = tp.ctgt[q];
        r.e = tp.eid ? tp.eid[q] : q;
    }
    return r;
}

// edges of local fibre lf inside the tile: [e0, e0 + n)
__device__ __forceinline__ void fibre_range(const Topo& tp, const Tile& t, int lf, int& e0, int& n) {
    if (tp.layout == PFS_LAYOUT_DENSE) {
        e0 = lf * tp.T;
        n = tp.T;
    } else {
        const int a = tp.rowptr[t.fibre0 + lf];
        e0 = a - t.q0;
        n = tp.rowptr[t.fibre0 + lf + 1] - a;
    }
}

// ------------------------------------------------------------------------------------------
// row loads / stores (rows of F, 2F, 4F floats; F is even so rows are at least 8-byte aligned)
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&x)[N]) {
    if constexpr (N % 4 == 0) {
        const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 v = __ldg(q + i);
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
    } else {
        const float2* q = reinterpret_cast<const float2*>(p);
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float2 v = __ldg(q + i);
            x[2 * i] = v.x; x[2 * i + 1] = v.y;
        }
    }
}

template <int N>
__device__ __forceinline__ void add_row(const float* __restrict__ p, float (&x)[N]) {
    if constexpr (N % 4 == 0) {
        const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 v = __ldg(q + i);
            x[4 * i] += v.x; x[4 * i + 1] += v.y; x[4 * i + 2] += v.z; x[4 * i + 3] += v.w;
        }
    } else {
        const float2* q = reinterpret_cast<const float2*>(p);
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float2 v = __ldg(q + i);
            x[2 * i] += v.x; x[2 * i + 1] += v.y;
        }
    }
}

template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&x)[N]) {
    if constexpr (N % 4 == 0) {
        float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; ++i) q[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    } else {
        float2* q = reinterpret_cast<float2*>(p);
#pragma unroll
        for (int i = 0; i < N / 2; ++i) q[i] = make_float2(x[2 * i], x[2 * i + 1]);
    }
}

// shared-memory row store with an arbitrary (even) leading dimension
template <int N>
__device__ __forceinline__ void store_row_smem(float* p, const float (&x)[N]) {
    float2* q = reinterpret_cast<float2*>(p);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) q[i] = make_float2(x[2 * i], x[2 * i + 1]);
}

// ------------------------------------------------------------------------------------------
// weights: global (torch Linear layout W[j][k], leading dimension ld) -> shared memory
// ------------------------------------------------------------------------------------------
// dst[k * J + j] = W[j * ld + koff + k]      ("input-major": forward layers, y_j += W_jk x_k)
template <int K, int J>
__device__ __forceinline__ void load_w_inmajor(float* dst, const float* __restrict__ W, int ld, int koff) {
    for (int i = threadIdx.x; i < J * K; i += blockDim.x) {
        const int j = i / K, k = i - j * K;
        dst[k * J + j] = __ldg(W + (size_t)j * ld + koff + k);
    }
}
// dst[j * K + k] = W[j * ld + koff + k]      ("output-major": backward layers, dx_k += W_jk dy_j)
template <int K, int J>
__device__ __forceinline__ void load_w_outmajor(float* dst, const float* __restrict__ W, int ld, int koff) {
    for (int i = threadIdx.x; i < J * K; i += blockDim.x) {
        const int j = i / K, k = i - j * K;
        dst[j * K + k] = __ldg(W + (size_t)j * ld + koff + k);
    }
}
template <int N>
__device__ __forceinline__ void load_vec(float* dst, const float* __restrict__ v) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) dst[i] = __ldg(v + i);
}

// y[j] += sum_k Wt[k * J + j] * x[k]; Wt in shared memory (every lane reads the same address:
// one broadcast wavefront per LDS.128), x and y in registers.
template <int K, int J>
__device__ __forceinline__ void dense_acc(const float* Wt, const float (&x)[K], float (&y)[J]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float xk = x[k];
        if constexpr (J % 4 == 0) {
            const float4* w = reinterpret_cast<const float4*>(Wt + k * J);
#pragma unroll
            for (int j = 0; j < J / 4; ++j) {
                const float4 v = w[j];
                y[4 * j] = fmaf(v.x, xk, y[4 * j]);
                y[4 * j + 1] = fmaf(v.y, xk, y[4 * j + 1]);
                y[4 * j + 2] = fmaf(v.z, xk, y[4 * j + 2]);
                y[4 * j + 3] = fmaf(v.w, xk, y[4 * j + 3]);
            }
        } else {
            const float2* w = reinterpret_cast<const float2*>(Wt + k * J);
#pragma unroll
            for (int j = 0; j < J / 2; ++j) {
                const float2 v = w[j];
                y[2 * j] = fmaf(v.x, xk, y[2 * j]);
                y[2 * j + 1] = fmaf(v.y, xk, y[2 * j + 1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// warp / block reductions (fixed shuffle tree => deterministic)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// BatchNorm tile statistics, two-pass inside the tile (robust against |mean| >> std):
// every thread holds one row z[N] (inactive threads contribute nothing); writes
// out[0..N) = tile mean, out[N..2N) = sum of squared deviations, out[2N] = count.
// `red` is shared scratch of (kWarps + 1) * N floats.
template <int N>
__device__ __forceinline__ void tile_bn_partial(const float (&z)[N], bool active, int count, float* red,
                                                float* __restrict__ out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* mean_s = red + kWarps * N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float v = warp_sum(active ? z[j] : 0.f);
        if (lane == 0) red[w * N + j] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) s += red[i * N + threadIdx.x];
        mean_s[threadIdx.x] = count > 0 ? s / (float)count : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float d = active ? z[j] - mean_s[j] : 0.f;
        const float v = warp_sum(d * d);
        if (lane == 0) red[w * N + j] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) s += red[i * N + threadIdx.x];
        out[threadIdx.x] = mean_s[threadIdx.x];
        out[N + threadIdx.x] = s;
        if (threadIdx.x == 0) out[2 * N] = (float)count;
    }
    __syncthreads();
}
__host__ __device__ constexpr int bn_partial_stride(int N) { return 2 * N + 2; }

// ------------------------------------------------------------------------------------------
// outer-product accumulation  dW[J][K] += sum_r D[r][0..J) (x) X[r][0..K)
// D and X are shared-memory tiles with leading dimensions ldD / ldX.  The J x K result is
// register-blocked TJ x TK per thread; rows are split over `groups` thread groups and the
// groups are summed in a fixed order by flush().
// ------------------------------------------------------------------------------------------
template <int J, int K, int TJ, int TK, int T0 = 0, int NT = kThreads>
struct OuterAcc {
    // threads [T0, T0 + NT) of the CTA take part; the others hold no accumulators in use
    static_assert(J % TJ == 0 && K % TK == 0, "block must divide the matrix");
    static constexpr int NB = (J / TJ) * (K / TK);
    static_assert(NB <= NT, "too many blocks for the thread range");
    static constexpr int GROUPS = NT / NB;
    static constexpr int kScratchFloats = GROUPS * J * K;
    float acc[TJ][TK];
    int grp, j0, k0;
    bool live;

    __device__ __forceinline__ void init() {
        const int t = (int)threadIdx.x - T0;
        grp = t >= 0 ? t / NB : GROUPS;
        live = t >= 0 && grp < GROUPS;
        const int b = t >= 0 ? t - grp * NB : 0;
        j0 = (b / (K / TK)) * TJ;
        k0 = (b % (K / TK)) * TK;
#pragma unroll
        for (int a = 0; a < TJ; ++a)
#pragma unroll
            for (int c = 0; c < TK; ++c) acc[a][c] = 0.f;
    }
    __device__ __forceinline__ void accumulate(const float* D, int ldD, const float* X, int ldX, int rows) {
        if (!live) return;
        for (int r = grp; r < rows; r += GROUPS) {
            float d[TJ], x[TK];
            const float* dp = D + r * ldD + j0;
            const float* xp = X + r * ldX + k0;
#pragma unroll
            for (int a = 0; a < TJ; ++a) d[a] = dp[a];
#pragma unroll
            for (int c = 0; c < TK; ++c) x[c] = xp[c];
#pragma unroll
            for (int a = 0; a < TJ; ++a)
#pragma unroll
                for (int c = 0; c < TK; ++c) acc[a][c] = fmaf(d[a], x[c], acc[a][c]);
        }
    }
    // Sums the groups through `scratch` (kScratchFloats floats of shared memory, may alias the
    // tiles) and writes the CTA's partial to out[j * ldo + ko + k].  Call with ALL threads of the CTA.
    __device__ __forceinline__ void flush(float* scratch, float* __restrict__ out, int ldo, int ko) {
        __syncthreads();
        if (live) {
#pragma unroll
            for (int a = 0; a < TJ; ++a)
#pragma unroll
                for (int c = 0; c < TK; ++c) scratch[grp * (J * K) + (j0 + a) * K + k0 + c] = acc[a][c];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < J * K; i += kThreads) {
            float s = 0.f;
            for (int g = 0; g < GROUPS; ++g) s += scratch[g * (J * K) + i];
            const int j = i / K, k = i - j * K;
            out[j * ldo + ko + k] = s;
        }
        __syncthreads();
    }
};

}  // namespace pfs
