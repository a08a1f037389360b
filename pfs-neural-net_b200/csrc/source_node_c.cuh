// source_node_c.cuh -- SModel node pass (reference src/gnn.py:147-154) with the MLP weights in the
// constant bank (Fdim <= 10: W3 [10F,9F] + W4 [F,10F] fit its 60 KB).
//
// Mapping: a CTA of 10 warps owns a tile of 128 fibres.  Warp w = (row set rs = w / 5, chunk c = w % 5):
// its lanes hold two fibres each (rows rs*64 + lane and + 32) and it produces the 2F hidden units of
// chunk c for them.  All lanes of a warp use the same weights, so the weights are warp-uniform
// operands: LDCU from the constant bank -> uniform registers -> FFMA2 R, R.F32, UR.F32x2, R.  The
// shared-memory return path only carries the activations (one row value per lane and k), not the
// weights, which is what capped the shared-memory version at ~25 % of the FMA pipe.
#pragma once
#include "source_model.cuh"

namespace pfs {

constexpr int kNodeThreadsC = 320;   // 10 warps = 2 row sets x 5 chunks
constexpr int kNodeRowsC = 128;      // fibres per node tile (2 per lane)

template <int F>
struct SourceNodeConst {
    static constexpr int K9 = 9 * F, J = 10 * F;
    // The node-MLP constants sit ABOVE the message-MLP constants (MsgEdgeConst, at most 2 * 16 * 32 + 2 * 32 * 32 + 32
    // floats at Fdim 16) so that one upload per module call serves the edge kernel and the node kernel.
    static constexpr int kBase = 4096;
    // forward: W4 input-major [J][F], b4 [F], then (FMA fallback only) W3 input-major [K9][J]
    static constexpr int kW4t = kBase, kB4 = kBase + J * F, kW3t = kBase + J * F + F, kFwdMmaFloats = kBase + J * F + F,
                         kFwdFloats = kBase + J * F + F + K9 * J;
    // backward: W4 as stored [F][J] (dh3_j += W4[f][j] dy_f), then (FMA fallback only) W3 as stored, first 9F columns [J][K9]
    static constexpr int kW4o = kBase, kW3o = kBase + F * J, kBwdMmaFloats = kBase + F * J, kBwdFloats = kBase + F * J + J * K9;
    static constexpr bool fits = kFwdFloats <= kConstFloats && kBwdFloats <= kConstFloats;
};

__device__ __forceinline__ int warp_index_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// hcat rows with any block size (same math as build_hcat)
template <int F>
__device__ __forceinline__ void build_hcat_n(const float* __restrict__ x_s, const float* __restrict__ moments,
                                             size_t row0, int rows, float* HC, int ld) {
    constexpr int M = 2 * F;
    for (int i = threadIdx.x; i < rows * F; i += blockDim.x) {
        const int r = i / F, k = i - r * F;
        HC[r * ld + k] = __ldg(x_s + (row0 + r) * F + k);
    }
    for (int i = threadIdx.x; i < rows * M; i += blockDim.x) {
        const int r = i / M, j = i - r * M;
        const float* mo = moments + (row0 + r) * 5 * M + j;
        const float mean = __ldg(mo), ex2 = __ldg(mo + M), c3 = __ldg(mo + 3 * M), c4 = __ldg(mo + 4 * M);
        const float vr = ex2 - mean * mean;
        const float var = vr > 0.f ? vr : kSlopeVar * vr;
        const float std0 = sqrtf(var + kStdEps);
        const float skew = c3 / (std0 * std0 * std0);
        const float kurt = c4 / (std0 * std0 * std0 * std0);
        float* h = HC + r * ld + F + j;
        h[0] = nan_to_num(mean);
        h[M] = sqrtf(nan_to_num(var) + kStdEps);
        h[2 * M] = nan_to_num(skew);
        h[3 * M] = nan_to_num(kurt);
    }
}

template <int F>
struct SourceNodeFwdSmemC {
    static constexpr int K9 = 9 * F, J = 10 * F;
    static constexpr int LDH = K9 + 1, LDA = J + 1;
    static constexpr int kFloats = kNodeRowsC * LDA + J + kNodeRowsC * F;
    static constexpr size_t bytes = sizeof(float) * kFloats;
};

template <int F>
__global__ void __launch_bounds__(kNodeThreadsC) k_source_node_fwd_c(const SourceNodeFwdParams p) {
    using SM = SourceNodeFwdSmemC<F>;
    using CW = SourceNodeConst<F>;
    constexpr int K9 = SM::K9, J = SM::J, C = 2 * F, LDH = SM::LDH, LDA = SM::LDA;
    extern __shared__ __align__(16) float sm[];
    float* BUF = sm;                        // HC [rows][LDH], then A3 [rows][LDA]
    float* b3e = BUF + kNodeRowsC * LDA;    // [J]  b3 + W3[:, 9F:] . u[g]
    float* YS = b3e + J;                    // [rows][F]
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const int chunk = warp % 5, rs = warp / 5;
    const int off = chunk * C;
    const int total = p.ntiles * p.G;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int g = tile / p.ntiles, lt = tile - g * p.ntiles;
        const int f0 = lt * kNodeRowsC;
        const int rows = min(kNodeRowsC, p.S - f0);
        const size_t row0 = (size_t)g * p.S + f0;
        build_hcat_n<F>(p.x_s, p.moments, row0, rows, BUF, LDH);
        for (int j = threadIdx.x; j < J; j += blockDim.x) {
            float s = __ldg(p.b3 + j);
            for (int k = 0; k < F; ++k) s = fmaf(__ldg(p.w3 + (size_t)j * J + K9 + k), __ldg(p.u + (size_t)g * F + k), s);
            b3e[j] = s;
        }
        __syncthreads();
        const int r0 = rs * 64 + lane, r1 = r0 + 32;
        float a0[C], a1[C];
        {
            const float* h0 = BUF + (r0 < rows ? r0 : 0) * LDH;
            const float* h1 = BUF + (r1 < rows ? r1 : 0) * LDH;
#pragma unroll
            for (int c = 0; c < C; ++c) a0[c] = a1[c] = b3e[off + c];
#pragma unroll 2
            for (int k = 0; k < K9; ++k) {
                const float2 x0 = make_float2(h0[k], h0[k]), x1 = make_float2(h1[k], h1[k]);
#pragma unroll
                for (int c = 0; c < C; c += 2) {
                    const float2 w = make_float2(c_w[CW::kW3t + off + k * J + c], c_w[CW::kW3t + off + k * J + c + 1]);
                    const float2 v0 = __ffma2_rn(w, x0, make_float2(a0[c], a0[c + 1]));
                    const float2 v1 = __ffma2_rn(w, x1, make_float2(a1[c], a1[c + 1]));
                    a0[c] = v0.x; a0[c + 1] = v0.y;
                    a1[c] = v1.x; a1[c + 1] = v1.y;
                }
            }
        }
        __syncthreads();   // every warp is done reading HC: reuse the buffer for the activations
        if (r0 < rows) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                a0[c] = lrelu(a0[c]);
                BUF[r0 * LDA + off + c] = a0[c];
            }
            store_row<C>(p.hidden + (row0 + r0) * J + off, a0);
        }
        if (r1 < rows) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                a1[c] = lrelu(a1[c]);
                BUF[r1 * LDA + off + c] = a1[c];
            }
            store_row<C>(p.hidden + (row0 + r1) * J + off, a1);
        }
        __syncthreads();
        // second layer, thread-per-row (warps 0..3), weights again warp-uniform constants
        if (threadIdx.x < rows) {
            const int r = threadIdx.x;
            float y[F];
#pragma unroll
            for (int f = 0; f < F; ++f) y[f] = c_w[CW::kB4 + f];
            const float* a = BUF + r * LDA;
#pragma unroll 4
            for (int k = 0; k < J; ++k) {
                const float2 xx = make_float2(a[k], a[k]);
#pragma unroll
                for (int f = 0; f < F; f += 2) {
                    const float2 v = __ffma2_rn(make_float2(c_w[CW::kW4t + k * F + f], c_w[CW::kW4t + k * F + f + 1]), xx,
                                                make_float2(y[f], y[f + 1]));
                    y[f] = v.x; y[f + 1] = v.y;
                }
            }
#pragma unroll
            for (int f = 0; f < F; ++f) YS[r * F + f] = y[f];
            store_row<F>(p.y_pre + (row0 + r) * F, y);
        }
        __syncthreads();
        if (p.bn_partial && threadIdx.x < F) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += YS[r * F + threadIdx.x];
            const float mean = s / (float)rows;
            float m2 = 0.f;
            for (int r = 0; r < rows; ++r) {
                const float d = YS[r * F + threadIdx.x] - mean;
                m2 += d * d;
            }
            float* o = p.bn_partial + (size_t)tile * bn_partial_stride(F);
            o[threadIdx.x] = mean;
            o[F + threadIdx.x] = m2;
            if (threadIdx.x == 0) o[2 * F] = (float)rows;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int F>
struct SourceNodeBwdSmemC {
    static constexpr int K9 = 9 * F, J = 10 * F;
    static constexpr int LDH = K9 + 1, LDA = J + 1, LDY = F + 1;
    // weight-gradient accumulators: dW3 on threads [0, 224) (or all 256 when F*F > 224), dW4 on one warp
    static constexpr int kNT3 = (F * F <= 224) ? 224 : kThreads;
    static constexpr int kT04 = (F * F <= 224) ? 224 : 0;
    static constexpr int kNT4 = ((F / 2) * 5 <= 32) ? 32 : 64;
    using AccW3 = OuterAcc<J, K9, 10, 9, 0, kNT3>;
    using AccW4 = OuterAcc<F, J, 2, 2 * F, kT04, kNT4>;
    static constexpr int kRows = kNodeRowsC * (LDH + 2 * LDA + LDY);
    static constexpr int kScr = AccW3::kScratchFloats > AccW4::kScratchFloats ? AccW3::kScratchFloats : AccW4::kScratchFloats;
    static constexpr int kFloats = kRows > kScr ? kRows : kScr;
    static constexpr size_t bytes = sizeof(float) * kFloats;
};

template <int F>
__global__ void __launch_bounds__(kNodeThreadsC) k_source_node_bwd_c(const SourceNodeBwdParams p) {
    using SM = SourceNodeBwdSmemC<F>;
    using CW = SourceNodeConst<F>;
    constexpr int K9 = SM::K9, J = SM::J, C = 2 * F, M = 2 * F, LDH = SM::LDH, LDA = SM::LDA, LDY = SM::LDY;
    extern __shared__ __align__(16) float sm[];
    float* HC = sm;                          // [rows][LDH]  hcat, later dhcat
    float* A3 = HC + kNodeRowsC * LDH;       // [rows][LDA]
    float* DH3 = A3 + kNodeRowsC * LDA;      // [rows][LDA]
    float* DY = DH3 + kNodeRowsC * LDA;      // [rows][LDY]
    typename SM::AccW3 accw3;
    typename SM::AccW4 accw4;
    accw3.init();
    accw4.init();
    float db4 = 0.f;   // thread f < F
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const int chunk = warp % 5, rs = warp / 5;
    const int off = chunk * C;
    const int total = p.ntiles * p.G;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int g = tile / p.ntiles, lt = tile - g * p.ntiles;
        const int f0 = lt * kNodeRowsC;
        const int rows = min(kNodeRowsC, p.S - f0);
        const size_t row0 = (size_t)g * p.S + f0;
        build_hcat_n<F>(p.x_s, p.moments, row0, rows, HC, LDH);
        // dy = BatchNorm backward of the upstream gradient
        {
            const float* sv = p.bn_save + (size_t)g * 4 * F;
            const float* st = p.bn_stat + (size_t)g * 2 * F;
            const float invS = 1.f / (float)p.S;
            for (int i = threadIdx.x; i < rows * F; i += blockDim.x) {
                const int r = i / F, f = i - r * F;
                const float gv = __ldg(p.gout + (row0 + r) * F + f);
                float dy;
                if (p.mode == 1) {
                    const float rstd = rsqrtf(sv[F + f] + p.eps);
                    const float xh = (__ldg(p.y_pre + (row0 + r) * F + f) - sv[f]) * rstd;
                    dy = sv[2 * F + f] * (gv - st[f] * invS - xh * st[F + f] * invS);
                } else if (p.mode == 2) {
                    dy = gv * sv[2 * F + f];
                } else {
                    dy = gv;
                }
                DY[r * LDY + f] = dy;
            }
        }
        __syncthreads();
        // dh3 = (dy . W4) * lrelu'(h3): two rows per lane, one chunk per warp
        {
            const int r0 = rs * 64 + lane, r1 = r0 + 32;
            const int q0 = r0 < rows ? r0 : 0, q1 = r1 < rows ? r1 : 0;
            float a0[C], a1[C], d0[C], d1[C];
            load_row<C>(p.hidden + (row0 + q0) * J + off, a0);
            load_row<C>(p.hidden + (row0 + q1) * J + off, a1);
#pragma unroll
            for (int c = 0; c < C; ++c) d0[c] = d1[c] = 0.f;
            const float* y0 = DY + q0 * LDY;
            const float* y1 = DY + q1 * LDY;
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float2 v0 = make_float2(y0[f], y0[f]), v1 = make_float2(y1[f], y1[f]);
#pragma unroll
                for (int c = 0; c < C; c += 2) {
                    const float2 w = make_float2(c_w[CW::kW4o + off + f * J + c], c_w[CW::kW4o + off + f * J + c + 1]);
                    const float2 e0 = __ffma2_rn(w, v0, make_float2(d0[c], d0[c + 1]));
                    const float2 e1 = __ffma2_rn(w, v1, make_float2(d1[c], d1[c + 1]));
                    d0[c] = e0.x; d0[c + 1] = e0.y;
                    d1[c] = e1.x; d1[c + 1] = e1.y;
                }
            }
            if (r0 < rows) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    A3[r0 * LDA + off + c] = a0[c];
                    DH3[r0 * LDA + off + c] = d0[c] * (a0[c] > 0.f ? 1.f : kSlope);
                }
            }
            if (r1 < rows) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    A3[r1 * LDA + off + c] = a1[c];
                    DH3[r1 * LDA + off + c] = d1[c] * (a1[c] > 0.f ? 1.f : kSlope);
                }
            }
        }
        __syncthreads();
        accw3.accumulate(DH3, LDA, HC, LDH, rows);
        accw4.accumulate(DY, LDY, A3, LDA, rows);
        if (threadIdx.x >= kThreads && threadIdx.x < kThreads + F) {   // warps 8-9 hold no accumulators
            const int f = threadIdx.x - kThreads;
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += DY[r * LDY + f];
            db4 += s;
        }
        __syncthreads();   // HC is dead from here: it receives dhcat
        // dhcat[r][k] = sum_j dh3[r][j] W3[j][k]: chunks of F columns, warp (rs, chunk) takes chunk and chunk + 5
        {
            const int r0 = rs * 64 + lane, r1 = r0 + 32;
            const int q0 = r0 < rows ? r0 : 0, q1 = r1 < rows ? r1 : 0;
            const float* d0 = DH3 + q0 * LDA;
            const float* d1 = DH3 + q1 * LDA;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int kc = chunk + 5 * pass;          // warp-uniform
                if (kc >= 9) break;
                const int koff = kc * F;
                float o0[F], o1[F];
#pragma unroll
                for (int k = 0; k < F; ++k) o0[k] = o1[k] = 0.f;
#pragma unroll 2
                for (int j = 0; j < J; ++j) {
                    const float2 v0 = make_float2(d0[j], d0[j]), v1 = make_float2(d1[j], d1[j]);
#pragma unroll
                    for (int k = 0; k < F; k += 2) {
                        const float2 w = make_float2(c_w[CW::kW3o + koff + j * K9 + k], c_w[CW::kW3o + koff + j * K9 + k + 1]);
                        const float2 e0 = __ffma2_rn(w, v0, make_float2(o0[k], o0[k + 1]));
                        const float2 e1 = __ffma2_rn(w, v1, make_float2(o1[k], o1[k + 1]));
                        o0[k] = e0.x; o0[k + 1] = e0.y;
                        o1[k] = e1.x; o1[k + 1] = e1.y;
                    }
                }
                if (r0 < rows) {
#pragma unroll
                    for (int k = 0; k < F; ++k) HC[r0 * LDH + koff + k] = o0[k];
                }
                if (r1 < rows) {
#pragma unroll
                    for (int k = 0; k < F; ++k) HC[r1 * LDH + koff + k] = o1[k];
                }
            }
        }
        __syncthreads();
        // direct gradient of x_s, column sums of dh3, moment-polynomial coefficients
        for (int i = threadIdx.x; i < rows * F; i += blockDim.x) {
            const int r = i / F, k = i - r * F;
            p.g_x_s[(row0 + r) * F + k] = HC[r * LDH + k];
        }
        for (int j = threadIdx.x; j < J; j += blockDim.x) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += DH3[r * LDA + j];
            p.tot3_part[(size_t)tile * J + j] = s;
        }
        for (int i = threadIdx.x; i < rows * M; i += blockDim.x) {
            const int r = i / M, j = i - r * M;
            const float* mo = p.moments + (row0 + r) * 5 * M + j;
            const float mean = __ldg(mo), ex2 = __ldg(mo + M), c2 = __ldg(mo + 2 * M), c3 = __ldg(mo + 3 * M),
                        c4 = __ldg(mo + 4 * M);
            const float* dh = HC + r * LDH + F + j;
            float d_mean = dh[0], d_std = dh[M], d_skew = dh[2 * M], d_kurt = dh[3 * M];
            const float vr = ex2 - mean * mean;
            const float var = vr > 0.f ? vr : kSlopeVar * vr;
            const float std0 = sqrtf(var + kStdEps);
            const float s3 = std0 * std0 * std0, s4 = s3 * std0;
            // torch: nan_to_num backward passes the gradient only where the value was finite
            if (!finite_f(mean)) d_mean = 0.f;
            if (!finite_f(var)) d_std = 0.f;
            if (!finite_f(c3 / s3)) d_skew = 0.f;
            if (!finite_f(c4 / s4)) d_kurt = 0.f;
            const float std1 = sqrtf(nan_to_num(var) + kStdEps);
            const float d_c3 = d_skew / s3, d_c4 = d_kurt / s4;
            const float d_var = d_std / (2.f * std1) + (-3.f * c3 / s4 * d_skew - 4.f * c4 / (s4 * std0) * d_kurt) / (2.f * std0);
            const float d_vr = d_var * (vr > 0.f ? 1.f : kSlopeVar);
            const float d_mu = d_mean - 2.f * mean * d_vr - 3.f * c2 * d_c3 - 4.f * c3 * d_c4;
            float* o = p.coefA + (row0 + r) * 4 * M + j;
            o[0] = d_mu;
            o[M] = 2.f * d_vr;
            o[2 * M] = 3.f * d_c3;
            o[3 * M] = 4.f * d_c4;
        }
        __syncthreads();
    }
    float* out = p.wpartial + (size_t)blockIdx.x * p.pstride;
    accw3.flush(HC, out, K9, 0);
    accw4.flush(HC, out + J * K9, J, 0);
    if (threadIdx.x >= kThreads && threadIdx.x < kThreads + F) out[J * K9 + F * J + threadIdx.x - kThreads] = db4;
}

}  // namespace pfs
