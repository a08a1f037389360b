// edge_model.cuh -- EdgeModel kernels (reference src/gnn.py:73-101).
//
// forward : z = W2 . lrelu(P_s[src] + P_t[tgt] + W1_e . x_e) + b2 per edge (thread-per-edge), with the
//           BatchNorm tile statistics of z; the double BatchNorm is then one per-feature affine.
// backward: recompute the hidden layer, closed-form double-BatchNorm backward, input gradients in
//           registers, weight gradients as a CTA-wide outer-product accumulation over the tile,
//           fibre sums / class sums of dh for the node tables.
#pragma once
#include "common.cuh"

namespace pfs {

struct EdgeFwdParams {
    Topo tp;
    const float* x_e;   // [G,E,F]
    const float* Ps;    // [G,S,4F]
    const float* Pt;    // [G,T,4F]  (includes W1_u.u + b1)
    const float* w1;    // [4F,4F]
    const float* w2;    // [F,4F]
    const float* b2;    // [F]
    float* z_out;       // [G,E,F]
    float* bn_partial;  // [G,ntiles,2F+2] or null
};

template <int F>
__global__ void __launch_bounds__(kThreads) k_edge_fwd(const EdgeFwdParams p) {
    constexpr int H = 4 * F;
    __shared__ __align__(16) float W1t[F * H];   // [k<F][j<H]
    __shared__ __align__(16) float W2t[H * F];   // [k<H][j<F]
    __shared__ float b2s[F];
    __shared__ float red[(kWarps + 1) * F];
    load_w_inmajor<F, H>(W1t, p.w1, H, 2 * F);
    load_w_inmajor<H, F>(W2t, p.w2, H, 0);
    load_vec<F>(b2s, p.b2);
    __syncthreads();
    const Topo& tp = p.tp;
    const int total = tp.ntiles * tp.G;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const Tile t = get_tile(tp, tile);
        const bool active = threadIdx.x < t.ne;
        float z[F];
#pragma unroll
        for (int j = 0; j < F; ++j) z[j] = 0.f;
        if (active) {
            const EdgeRef er = get_edge(tp, t, threadIdx.x);
            float x[F], h[H];
            load_row<F>(p.x_e + ((size_t)t.g * tp.E + er.e) * F, x);
            load_row<H>(p.Ps + ((size_t)t.g * tp.S + er.src) * H, h);
            add_row<H>(p.Pt + ((size_t)t.g * tp.T + er.tgt) * H, h);
            dense_acc<F, H>(W1t, x, h);
#pragma unroll
            for (int j = 0; j < H; ++j) h[j] = lrelu(h[j]);
#pragma unroll
            for (int j = 0; j < F; ++j) z[j] = b2s[j];
            dense_acc<H, F>(W2t, h, z);
            store_row<F>(p.z_out + ((size_t)t.g * tp.E + er.e) * F, z);
        }
        if (p.bn_partial)
            tile_bn_partial<F>(z, active, t.ne, red, p.bn_partial + (size_t)tile * bn_partial_stride(F));
    }
}

// ------------------------------------------------------------------------------------------
// backward statistics of the (double) BatchNorm over the edges of each graph:
// partial[tile][0..F) = sum g, partial[tile][F..2F) = sum g * xhat1, xhat1 = (x_e_out - beta) * inv
// ------------------------------------------------------------------------------------------
struct EdgeBnStatParams {
    Topo tp;
    const float* xe2;    // [G,E,F] forward output
    const float* gout;   // [G,E,F]
    const float* coef;   // [G,6,F]: A, c1, c2, inv, beta, (unused)   -- only inv and beta are read here
    float* partial;      // [G*ntiles][2F]
};

template <int F>
__global__ void __launch_bounds__(kThreads) k_edge_bn_bwd_stats(const EdgeBnStatParams p) {
    __shared__ float red[kWarps * 2 * F];
    const Topo& tp = p.tp;
    const int total = tp.ntiles * tp.G;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const Tile t = get_tile(tp, tile);
        const bool active = threadIdx.x < t.ne;
        float g[F], xh[F];
#pragma unroll
        for (int j = 0; j < F; ++j) g[j] = xh[j] = 0.f;
        if (active) {
            const EdgeRef er = get_edge(tp, t, threadIdx.x);
            const size_t row = ((size_t)t.g * tp.E + er.e) * F;
            load_row<F>(p.gout + row, g);
            load_row<F>(p.xe2 + row, xh);
            const float* c = p.coef + (size_t)t.g * 6 * F;
#pragma unroll
            for (int j = 0; j < F; ++j) xh[j] = (xh[j] - c[4 * F + j]) * c[3 * F + j];
        }
#pragma unroll
        for (int j = 0; j < F; ++j) {
            const float a = warp_sum(g[j]);
            const float b = warp_sum(g[j] * xh[j]);
            if (lane == 0) {
                red[w * 2 * F + j] = a;
                red[w * 2 * F + F + j] = b;
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * F) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) s += red[i * 2 * F + threadIdx.x];
            p.partial[(size_t)tile * 2 * F + threadIdx.x] = s;
        }
        __syncthreads();
    }
}

// Per-graph coefficient vectors of the closed-form backward (tests/kernel_model.py: edge_bwd).
//   stage 0 (before the statistics pass): coef[g] = {A, 0, 0, inv, beta}
//   stage 1 (after it): fills c1 = mean g, c2 = mean(g xhat1) * kappa and the per-graph
//            gamma / beta gradients dgb[g][0..F) , dgb[g][F..2F)
// mode: 0 = not normed, 1 = train, 2 = eval
__global__ void k_edge_bn_bwd_coef(int stage, int mode, int F, int G, int ntiles, long long n_rows,
                                   const float* __restrict__ save, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ rm,
                                   const float* __restrict__ rv, float eps, const float* __restrict__ partial,
                                   float* __restrict__ coef, float* __restrict__ dgb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * F) return;
    const int g = i / F, f = i - g * F;
    float* c = coef + (size_t)g * 6 * F;
    if (mode == 0) {
        c[f] = 1.f; c[F + f] = 0.f; c[2 * F + f] = 0.f; c[3 * F + f] = 0.f; c[4 * F + f] = 0.f;
        return;
    }
    const double gm = gamma[f], bt = beta[f];
    const float* s = save + (size_t)g * 4 * F;
    if (mode == 2) {
        // eval: y1 = a (z - m) + bt, y2 = a (y1 - m) + bt, dz = a^2 g
        const double cc = 1.0 / sqrt((double)rv[f] + (double)eps), a = gm * cc, m = rm[f];
        if (stage == 0) {
            c[f] = (float)(a * a); c[F + f] = 0.f; c[2 * F + f] = 0.f;
            // xhat slot is reused to recover z: z = (xe2 - shift) / A
            c[3 * F + f] = (a != 0.0) ? (float)(1.0 / (a * a)) : 0.f;
            c[4 * F + f] = s[3 * F + f];
        } else {
            // sum g and sum g * z come from the statistics pass (xhat == z here)
            double sg = 0.0, sgz = 0.0;
            for (int t = 0; t < ntiles; ++t) {
                sg += partial[((size_t)g * ntiles + t) * 2 * F + f];
                sgz += partial[((size_t)g * ntiles + t) * 2 * F + F + f];
            }
            // dgamma = sum g c [(y1 - m) + a (z - m)] = c [ 2a (sgz - m sg) + (bt - m) sg ]
            dgb[(size_t)g * 2 * F + f] = (float)(cc * (2.0 * a * (sgz - m * sg) + (bt - m) * sg));
            dgb[(size_t)g * 2 * F + F + f] = (float)((a + 1.0) * sg);
        }
        return;
    }
    const double var = s[F + f];
    const double r1 = 1.0 / sqrt(var + (double)eps);
    const double var2 = gm * gm * var * r1 * r1;
    const double r2 = 1.0 / sqrt(var2 + (double)eps);
    if (stage == 0) {
        c[f] = (float)(gm * gm * r1 * r2);
        c[F + f] = 0.f;
        c[2 * F + f] = 0.f;
        c[3 * F + f] = (gm != 0.0) ? (float)(1.0 / (gm * gm * r2)) : 0.f;
        c[4 * F + f] = (float)bt;
        return;
    }
    double sg = 0.0, sgx = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        sg += partial[((size_t)g * ntiles + t) * 2 * F + f];
        sgx += partial[((size_t)g * ntiles + t) * 2 * F + F + f];
    }
    const double n = (double)n_rows;
    const double gbar = sg / n, mgx = sgx / n;
    const double sc = gm * r2, q = var * r1 * r1;
    const double kappa = sc * sc + 1.0 - sc * sc * q;
    c[F + f] = (float)gbar;
    c[2 * F + f] = (float)(mgx * kappa);
    dgb[(size_t)g * 2 * F + f] = (float)(n * mgx * sc * (2.0 - sc * sc * q));
    dgb[(size_t)g * 2 * F + F + f] = (float)(n * gbar);
}

struct EdgeBwdParams {
    Topo tp;
    const float *x_e, *xe2, *gout;   // [G,E,F]
    const float *Ps, *Pt;            // [G,S,4F], [G,T,4F]
    const float *w1, *w2;            // [4F,4F], [F,4F]
    const float* coef;               // [G,6,F]
    float* g_x_e;                    // [G,E,F]
    float* dPs;                      // [G,S,4F] fibre sums of dh
    float* class_part;               // dense: [G,ntiles,T,4F]
    float* dh_rows;                  // general: [G,E(q),4F]
    float* wpartial;                 // [ncta][pstride]: dW1_e [4F*F], dW2 [F*4F], db2 [F]
    int pstride;
};

template <int F>
struct EdgeBwdSmem {
    static constexpr int H = 4 * F;
    static constexpr int LDH = H + 4;   // 16-byte aligned rows, bank-staggered
    static constexpr int LDF = F + 2;
    static constexpr int kWeights = 3 * F * H;
    static constexpr int kTiles = kTile * (2 * LDH + 2 * LDF);
    static constexpr size_t bytes = sizeof(float) * (kWeights + kTiles);
};

template <int F>
__global__ void __launch_bounds__(kThreads) k_edge_bwd(const EdgeBwdParams p) {
    constexpr int H = 4 * F;
    using SM = EdgeBwdSmem<F>;
    constexpr int LDH = SM::LDH, LDF = SM::LDF;
    extern __shared__ __align__(16) float sm[];
    float* W1t = sm;                 // [k<F][j<H]   forward layer 1 (edge columns)
    float* W2o = W1t + F * H;        // [j<F][k<H]   backward through layer 2: da_k += W2[j][k] dz_j
    float* W1o = W2o + F * H;        // [j<H][k<F]   backward through layer 1: dx_k += W1[j][2F+k] dh_j
    float* DH = W1o + F * H;         // [kTile][LDH]
    float* A1 = DH + kTile * LDH;    // [kTile][LDH]
    float* DZ = A1 + kTile * LDH;    // [kTile][LDF]
    float* XE = DZ + kTile * LDF;    // [kTile][LDF]
    load_w_inmajor<F, H>(W1t, p.w1, H, 2 * F);
    load_w_outmajor<H, F>(W2o, p.w2, H, 0);
    load_w_outmajor<F, H>(W1o, p.w1, H, 2 * F);
    __syncthreads();
    // the two weight-gradient accumulations run side by side on the two halves of the CTA
    using AccW1 = OuterAcc<H, F, 8, F / 2, 0, kThreads / 2>;              // dW1_e[j][k] = sum dh_j x_k
    using AccW2 = OuterAcc<F, H, F / 2, 8, kThreads / 2, kThreads / 2>;   // dW2[j][k]   = sum dz_j a1_k
    AccW1 accw1;
    AccW2 accw2;
    accw1.init();
    accw2.init();
    float dzsum[F];
#pragma unroll
    for (int j = 0; j < F; ++j) dzsum[j] = 0.f;

    const Topo& tp = p.tp;
    const int total = tp.ntiles * tp.G;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const Tile t = get_tile(tp, tile);
        const bool active = threadIdx.x < t.ne;
        if (active) {
            const EdgeRef er = get_edge(tp, t, threadIdx.x);
            const size_t row = ((size_t)t.g * tp.E + er.e) * F;
            float x[F], h[H], dz[F];
            load_row<F>(p.x_e + row, x);
            load_row<H>(p.Ps + ((size_t)t.g * tp.S + er.src) * H, h);
            add_row<H>(p.Pt + ((size_t)t.g * tp.T + er.tgt) * H, h);
            dense_acc<F, H>(W1t, x, h);
            {
                float gr[F], xo[F];
                load_row<F>(p.gout + row, gr);
                load_row<F>(p.xe2 + row, xo);
                const float* c = p.coef + (size_t)t.g * 6 * F;
#pragma unroll
                for (int j = 0; j < F; ++j) {
                    const float xh = (xo[j] - c[4 * F + j]) * c[3 * F + j];
                    dz[j] = c[j] * (gr[j] - c[F + j] - xh * c[2 * F + j]);
                    dzsum[j] += dz[j];
                }
            }
            float da[H];
#pragma unroll
            for (int k = 0; k < H; ++k) da[k] = 0.f;
            dense_acc<F, H>(W2o, dz, da);
            store_row_smem<F>(DZ + threadIdx.x * LDF, dz);
            store_row_smem<F>(XE + threadIdx.x * LDF, x);
#pragma unroll
            for (int k = 0; k < H; ++k) {
                da[k] *= dlrelu(h[k]);      // dh
                h[k] = lrelu(h[k]);         // a1
            }
            store_row_smem<H>(A1 + threadIdx.x * LDH, h);
            store_row_smem<H>(DH + threadIdx.x * LDH, da);
            float dx[F];
#pragma unroll
            for (int k = 0; k < F; ++k) dx[k] = 0.f;
            dense_acc<H, F>(W1o, da, dx);
            store_row<F>(p.g_x_e + row, dx);
            if (p.dh_rows) store_row<H>(p.dh_rows + ((size_t)t.g * tp.E + t.q0 + threadIdx.x) * H, da);
        }
        __syncthreads();
        accw1.accumulate(DH, LDH, XE, LDF, t.ne);
        accw2.accumulate(DZ, LDF, A1, LDH, t.ne);
        // fibre sums of dh -> dPs
        for (int i = threadIdx.x; i < t.nfib * H; i += kThreads) {
            const int lf = i / H, k = i - lf * H;
            int e0, n;
            fibre_range(tp, t, lf, e0, n);
            float s = 0.f;
            for (int e = 0; e < n; ++e) s += DH[(e0 + e) * LDH + k];
            p.dPs[((size_t)t.g * tp.S + t.fibre0 + lf) * H + k] = s;
        }
        // class sums of dh (dense layout: edge lf*T + c belongs to class c)
        if (p.class_part) {
            float* cp = p.class_part + (size_t)tile * tp.T * H;
            for (int i = threadIdx.x; i < tp.T * H; i += kThreads) {
                const int c = i / H, k = i - c * H;
                float s = 0.f;
                for (int lf = 0; lf < t.nfib; ++lf) s += DH[(lf * tp.T + c) * LDH + k];
                cp[i] = s;
            }
        }
        __syncthreads();
    }
    float* out = p.wpartial + (size_t)blockIdx.x * p.pstride;
    accw1.flush(DH, out, F, 0);
    accw2.flush(DH, out + H * F, H, 0);
    // db2 = sum dz: block reduction of the per-thread sums
    {
        float* red = DH;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < F; ++j) {
            const float v = warp_sum(dzsum[j]);
            if (lane == 0) red[w * F + j] = v;
        }
        __syncthreads();
        if (threadIdx.x < F) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) s += red[i * F + threadIdx.x];
            out[2 * H * F + threadIdx.x] = s;
        }
    }
}

}  // namespace pfs
