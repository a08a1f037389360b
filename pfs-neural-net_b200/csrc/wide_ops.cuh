// wide_ops.cuh -- HBM-bound kernels of the wide-feature path (bf16 rows of 2F .. 10F features):
// column statistics (BatchNorm forward / backward sums, bias gradients), per-feature affine maps,
// deterministic segmented reductions over fibres and classes (reference torch_scatter call sites
// src/gnn.py:140-144,190 -- no atomics, fixed summation order), the moment statistics of SModel and
// their backward, and small layout helpers.  Every kernel moves rows as 4-byte bf16 pairs so a warp
// reads 128 contiguous bytes of a row.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfs {

using bf16 = __nv_bfloat16;
using bf162 = __nv_bfloat162;

__device__ __forceinline__ float2 ld_pair(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const bf162*>(p)); }
__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st_pair(bf16* p, float a, float b) { *reinterpret_cast<bf162*>(p) = __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ void st_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// 8 bf16 columns (16 bytes) per thread: the access width of the streaming kernels below
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(u[i] << 16);
        v[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
    uint4 w;
    bf162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    bf162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    w.x = *reinterpret_cast<uint32_t*>(&a); w.y = *reinterpret_cast<uint32_t*>(&b);
    w.z = *reinterpret_cast<uint32_t*>(&c); w.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = w;
}
__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void ld8_any(const bf16* p, float (&v)[8]) { ld8(p, v); }
__device__ __forceinline__ void ld8_any(const float* p, float (&v)[8]) { ld8f(p, v); }

// segments of rows: fibres or classes, dense (implicit) or listed
struct SegDesc {
    int mode;          // 0: dense fibre (rows seg*T + i, i < T), 1: dense class (rows i*T + seg, i < S), 2: list
    int nseg, S, T;
    const int* ptr;    // [nseg+1] (mode 2)
    const int* list;   // rows of segment seg: list[ptr[seg] .. ptr[seg+1])  (null: the positions themselves)
};
__device__ __forceinline__ int seg_len(const SegDesc& sd, int seg) {
    return sd.mode == 0 ? sd.T : sd.mode == 1 ? sd.S : sd.ptr[seg + 1] - sd.ptr[seg];
}
__device__ __forceinline__ long long seg_row(const SegDesc& sd, int seg, int i) {
    if (sd.mode == 0) return (long long)seg * sd.T + i;
    if (sd.mode == 1) return (long long)i * sd.T + seg;
    const int p = sd.ptr[seg] + i;
    return sd.list ? sd.list[p] : p;
}

// ------------------------------------------------------------------------------------------------
// column statistics over the rows of x [R, C] (leading dimension ld):
//   kind 0 (moments): s0 = sum (x - shift), s1 = sum (x - shift)^2, shift[c] = x[0][c]
//   kind 1 (BN backward): s0 = sum w g, s1 = sum w g (v - p0) p1    (v, p0, p1 optional: s1 = 0 without v)
// partial [nrb][2][C]; block (32, 8): x = column pair, y = row lane; grid (ceil(C / 64), nrb)
// ------------------------------------------------------------------------------------------------
template <class TG, class TV>
__global__ void __launch_bounds__(256) k_wide_colstats(int kind, const TG* __restrict__ g, int ldg, const TV* __restrict__ v,
                                                       int ldv, const float* __restrict__ p0, const float* __restrict__ p1,
                                                       const float* __restrict__ roww, long long R, int C,
                                                       long long rows_per_block, float* __restrict__ partial) {
    __shared__ float red[8][32][4];
    const int c = 2 * (blockIdx.x * 32 + threadIdx.x);
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(R, r0 + rows_per_block);
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    if (c < C) {
        float sh0 = 0.f, sh1 = 0.f, q0 = 0.f, q1 = 0.f, e0 = 1.f, e1 = 1.f;
        if (kind == 0) {
            const float2 s = ld_pair(g + c);
            sh0 = s.x; sh1 = s.y;
        } else if (v) {
            if (p0) { q0 = p0[c]; q1 = p0[c + 1]; }
            if (p1) { e0 = p1[c]; e1 = p1[c + 1]; }
        }
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
            const float2 x = ld_pair(g + r * ldg + c);
            if (kind == 0) {
                const float d0 = x.x - sh0, d1 = x.y - sh1;
                a0 += d0; a1 += d1;
                b0 = fmaf(d0, d0, b0); b1 = fmaf(d1, d1, b1);
            } else {
                const float w = roww ? roww[r] : 1.f;
                const float g0 = w * x.x, g1 = w * x.y;
                a0 += g0; a1 += g1;
                if (v) {
                    const float2 y = ld_pair(v + r * ldv + c);
                    b0 = fmaf(g0, (y.x - q0) * e0, b0);
                    b1 = fmaf(g1, (y.y - q1) * e1, b1);
                }
            }
        }
    }
    red[threadIdx.y][threadIdx.x][0] = a0;
    red[threadIdx.y][threadIdx.x][1] = a1;
    red[threadIdx.y][threadIdx.x][2] = b0;
    red[threadIdx.y][threadIdx.x][3] = b1;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int y = 0; y < 8; ++y)
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] += red[y][threadIdx.x][q];
        float* o = partial + (size_t)blockIdx.y * 2 * C;
        o[c] = s[0]; o[c + 1] = s[1];
        o[C + c] = s[2]; o[C + c + 1] = s[3];
    }
}
// out[0][c] = sum over row blocks of s0, out[1][c] = of s1 (fp64 accumulation, fixed order);
// kind 0 converts to (mean, M2) using the shift row
template <class TG>
__global__ void k_wide_colstats_final(int kind, const float* __restrict__ partial, int nrb, int C, long long R,
                                      const TG* __restrict__ shift_row, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    for (int b = 0; b < nrb; ++b) {
        s0 += partial[(size_t)b * 2 * C + c];
        s1 += partial[(size_t)b * 2 * C + C + c];
    }
    if (kind == 0) {
        const double n = (double)R;
        const double sh = (double)(float)shift_row[c];
        out[c] = (float)(sh + s0 / n);
        const double m2 = s1 - s0 * s0 / n;
        out[C + c] = (float)(m2 > 0.0 ? m2 : 0.0);
    } else {
        out[c] = (float)s0;
        out[C + c] = (float)s1;
    }
}

// ------------------------------------------------------------------------------------------------
// per-feature maps over rows
//   kind 0: out = a[c] * x + b[c]
//   kind 1: out = a[c] * (x - b[c] - (v - p0[c]) * p1[c] * c2[c])          (BatchNorm backward)
// ------------------------------------------------------------------------------------------------
template <class TX, class TV>
__global__ void __launch_bounds__(256) k_wide_rowmap(int kind, const TX* __restrict__ x, int ldx, const TV* __restrict__ v, int ldv,
                                                     const float* __restrict__ a, const float* __restrict__ b,
                                                     const float* __restrict__ p0, const float* __restrict__ p1,
                                                     const float* __restrict__ c2, long long R, int C,
                                                     bf16* __restrict__ out, int ldo) {
    const int half = C >> 1;
    const long long total = R * half;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / half;
        const int c = 2 * (int)(i - r * half);
        const float2 xx = ld_pair(x + r * ldx + c);
        float o0, o1;
        if (kind == 0) {
            o0 = fmaf(a[c], xx.x, b[c]);
            o1 = fmaf(a[c + 1], xx.y, b[c + 1]);
        } else {
            const float2 vv = ld_pair(v + r * ldv + c);
            o0 = a[c] * (xx.x - b[c] - (vv.x - p0[c]) * p1[c] * c2[c]);
            o1 = a[c + 1] * (xx.y - b[c + 1] - (vv.y - p0[c + 1]) * p1[c + 1] * c2[c + 1]);
        }
        st_pair(out + r * ldo + c, o0, o1);
    }
}

// the same maps with 8 columns per thread (16-byte accesses) and the per-column coefficients held in registers:
// a thread keeps its column group and strides over the rows
template <class TX, class TV>
__global__ void __launch_bounds__(256) k_wide_rowmap8(int kind, const TX* __restrict__ x, int ldx, const TV* __restrict__ v, int ldv,
                                                      const float* __restrict__ a, const float* __restrict__ b,
                                                      const float* __restrict__ p0, const float* __restrict__ p1,
                                                      const float* __restrict__ c2, long long R, int C,
                                                      bf16* __restrict__ out, int ldo) {
    const int groups = C >> 3;
    const int lanes = blockDim.x / groups;
    const int lr = threadIdx.x / groups, c = 8 * (threadIdx.x - lr * groups);
    if (lr >= lanes) return;
    float ca[8], cb[8], cs[8], cq[8];       // kind 1: out = ca (x - cb - (v - cq) cs), cs = p1 c2
    ld8f(a + c, ca);
    ld8f(b + c, cb);
    if (kind == 1) {
        float t1[8], t2[8];
        ld8f(p0 + c, cq);
        ld8f(p1 + c, t1);
        ld8f(c2 + c, t2);
#pragma unroll
        for (int q = 0; q < 8; ++q) cs[q] = t1[q] * t2[q];
    }
    for (long long r = (long long)blockIdx.x * lanes + lr; r < R; r += (long long)gridDim.x * lanes) {
        float xv[8], o[8];
        ld8_any(x + r * ldx + c, xv);
        if (kind == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = fmaf(ca[q], xv[q], cb[q]);
        } else {
            float vv[8];
            ld8_any(v + r * ldv + c, vv);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = ca[q] * (xv[q] - cb[q] - (vv[q] - cq[q]) * cs[q]);
        }
        st8(out + r * ldo + c, o);
    }
}

// ------------------------------------------------------------------------------------------------
// segmented sums: out[seg][c] = sum over the segment's rows of x[row][c]
// grid (nseg, nchunk), 256 threads, thread = column pair (loops when C > 512);
// nchunk > 1 writes partial[chunk][seg][C] for k_wide_segsum_final
// ------------------------------------------------------------------------------------------------
template <class TX>
__global__ void __launch_bounds__(256) k_wide_segsum(const SegDesc sd, const TX* __restrict__ x, int ldx, int C, int nchunk,
                                                     float* __restrict__ out_f32, bf16* __restrict__ out_bf16,
                                                     float* __restrict__ partial) {
    // thread = 8 columns (one 16-byte load) of one row lane; tpr threads cover a row, 256 / tpr rows are
    // in flight per pass and 4 passes are unrolled, so a block keeps 16 KB of loads outstanding
    __shared__ float red[256 * 8];
    const int seg = blockIdx.x, chunk = blockIdx.y;
    const int len = seg_len(sd, seg);
    const int i0 = (int)((long long)len * chunk / nchunk), i1 = (int)((long long)len * (chunk + 1) / nchunk);
    const int cgroups = C >> 3;
    const int tpr = cgroups < 256 ? cgroups : 256;           // threads per row
    const int lanes = 256 / tpr;                             // row lanes
    const int lc = threadIdx.x % tpr, lr = threadIdx.x / tpr;
    for (int cg0 = 0; cg0 < cgroups; cg0 += tpr) {
        const int cg = cg0 + lc;
        const bool live = lr < lanes && cg < cgroups;
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        if (live) {
            const TX* xc = x + 8 * cg;
            int i = i0 + lr;
            for (; i + 3 * lanes < i1; i += 4 * lanes) {
                float v0[8], v1[8], v2[8], v3[8];
                ld8_any(xc + seg_row(sd, seg, i) * ldx, v0);
                ld8_any(xc + seg_row(sd, seg, i + lanes) * ldx, v1);
                ld8_any(xc + seg_row(sd, seg, i + 2 * lanes) * ldx, v2);
                ld8_any(xc + seg_row(sd, seg, i + 3 * lanes) * ldx, v3);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += (v0[q] + v1[q]) + (v2[q] + v3[q]);
            }
            for (; i < i1; i += lanes) {
                float v0[8];
                ld8_any(xc + seg_row(sd, seg, i) * ldx, v0);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += v0[q];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) red[threadIdx.x * 8 + q] = acc[q];
        __syncthreads();
        if (lr == 0 && cg < cgroups) {
            float sum[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) sum[q] = 0.f;
            for (int r = 0; r < lanes; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) sum[q] += red[(r * tpr + lc) * 8 + q];
            const int c = 8 * cg;
            if (nchunk > 1) {
                float* p = partial + ((size_t)chunk * sd.nseg + seg) * C + c;
#pragma unroll
                for (int q = 0; q < 8; ++q) p[q] = sum[q];
            } else {
                if (out_f32) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) out_f32[(size_t)seg * C + c + q] = sum[q];
                }
                if (out_bf16) st8(out_bf16 + (size_t)seg * C + c, sum);
            }
        }
    }
}
__global__ void k_wide_segsum_final(const float* __restrict__ partial, int nchunk, long long n, float* __restrict__ out_f32,
                                    bf16* __restrict__ out_bf16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int q = 0; q < nchunk; ++q) s += partial[(size_t)q * n + i];
    if (out_f32) out_f32[i] = s;
    if (out_bf16) out_bf16[i] = __float2bfloat16_rn(s);
}

// ------------------------------------------------------------------------------------------------
// SModel moments (reference src/gnn.py:140-144): per fibre over its message rows m [E, C], C = 2F:
// moments[fibre] = {mean, E[m^2], c2, c3, c4} (central moments about the mean, two passes), [S,5,C]
// grid (S), C / 2 threads
// ------------------------------------------------------------------------------------------------
template <class TM>
__global__ void __launch_bounds__(256) k_wide_moments_fwd(const SegDesc sd, const TM* __restrict__ m, int C,
                                                          float* __restrict__ moments) {
    // same thread mapping as k_wide_segsum (8 columns x row lanes); pass 1: sum, sum of squares; pass 2: central moments
    __shared__ float red[256 * 8 * 3];
    __shared__ float mean_s[2048];
    const int seg = blockIdx.x;
    const int len = seg_len(sd, seg);
    const float inv = 1.f / (float)max(len, 1);
    const int cgroups = C >> 3;
    const int tpr = cgroups < 256 ? cgroups : 256;
    const int lanes = 256 / tpr;
    const int lc = threadIdx.x % tpr, lr = threadIdx.x / tpr;
    for (int cg0 = 0; cg0 < cgroups; cg0 += tpr) {
        const int cg = cg0 + lc;
        const bool live = lr < lanes && cg < cgroups;
        const TM* xc = m + 8 * cg;
        float s1[8], s2[8], s3[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) s1[q] = s2[q] = s3[q] = 0.f;
        if (live) {
            int i = lr;
            for (; i + lanes < len; i += 2 * lanes) {
                float v0[8], v1[8];
                ld8_any(xc + seg_row(sd, seg, i) * C, v0);
                ld8_any(xc + seg_row(sd, seg, i + lanes) * C, v1);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    s1[q] += v0[q] + v1[q];
                    s2[q] = fmaf(v0[q], v0[q], fmaf(v1[q], v1[q], s2[q]));
                }
            }
            for (; i < len; i += lanes) {
                float v0[8];
                ld8_any(xc + seg_row(sd, seg, i) * C, v0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    s1[q] += v0[q];
                    s2[q] = fmaf(v0[q], v0[q], s2[q]);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            red[threadIdx.x * 8 + q] = s1[q];
            red[2048 + threadIdx.x * 8 + q] = s2[q];
        }
        __syncthreads();
        float* o = moments + (size_t)seg * 5 * C + 8 * cg;
        if (lr == 0 && cg < cgroups) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float a = 0.f, b = 0.f;
                for (int r = 0; r < lanes; ++r) {
                    a += red[(r * tpr + lc) * 8 + q];
                    b += red[2048 + (r * tpr + lc) * 8 + q];
                }
                mean_s[lc * 8 + q] = a * inv;
                o[q] = a * inv;
                o[C + q] = b * inv;
            }
        }
        __syncthreads();
        float mean[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            mean[q] = mean_s[lc * 8 + q];
            s1[q] = s2[q] = s3[q] = 0.f;
        }
        if (live) {
            for (int i = lr; i < len; i += lanes) {
                float v0[8];
                ld8_any(xc + seg_row(sd, seg, i) * C, v0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float d = v0[q] - mean[q], d2 = d * d;
                    s1[q] += d2;
                    s2[q] = fmaf(d2, d, s2[q]);
                    s3[q] = fmaf(d2, d2, s3[q]);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            red[threadIdx.x * 8 + q] = s1[q];
            red[2048 + threadIdx.x * 8 + q] = s2[q];
            red[4096 + threadIdx.x * 8 + q] = s3[q];
        }
        __syncthreads();
        if (lr == 0 && cg < cgroups) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float a = 0.f, b = 0.f, c = 0.f;
                for (int r = 0; r < lanes; ++r) {
                    a += red[(r * tpr + lc) * 8 + q];
                    b += red[2048 + (r * tpr + lc) * 8 + q];
                    c += red[4096 + (r * tpr + lc) * 8 + q];
                }
                o[2 * C + q] = a * inv;
                o[3 * C + q] = b * inv;
                o[4 * C + q] = c * inv;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float wide_nan_to_num(float x) {   // torch.nan_to_num(nan=0): +-inf -> +-FLT_MAX
    if (x != x) return 0.f;
    if (isinf(x)) return x > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
    return x;
}
struct MomentStats {
    float mean, std, skew, kurt, var_raw, std0;
};
__device__ __forceinline__ MomentStats wide_moment_stats(float mean, float ex2, float c3, float c4) {
    MomentStats s;
    s.var_raw = ex2 - mean * mean;
    const float var = s.var_raw > 0.f ? s.var_raw : 0.01f * s.var_raw;    // F.leaky_relu default slope (src/gnn.py:141)
    s.std0 = sqrtf(var + 1e-6f);
    const float s3 = s.std0 * s.std0 * s.std0;
    s.mean = wide_nan_to_num(mean);
    s.std = sqrtf(wide_nan_to_num(var) + 1e-6f);
    s.skew = wide_nan_to_num(c3 / s3);
    s.kurt = wide_nan_to_num(c4 / (s3 * s.std0));
    return s;
}
// hcat[fibre] = [x_s | mean | std | skew | kurt]  (bf16 [S, 9F], row stride ldo; reference src/gnn.py:147-152).
// with_lo: the row continues with the bf16 remainders of the four statistics, [... | lo(mean) | lo(std) | lo(skew) |
// lo(kurt)] (17F columns): contracted against [W3 | W3[:, F:9F]] the fibre MLP sees the statistics to ~2^-17
__global__ void k_wide_source_hcat(const bf16* __restrict__ x_s, const float* __restrict__ moments, int S, int F,
                                   bf16* __restrict__ hcat, int ldo, int with_lo) {
    const int C = 2 * F, K9 = 9 * F;
    const long long total = (long long)S * (F + C);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / (F + C);
        const int j = (int)(i - s * (F + C));
        bf16* o = hcat + s * ldo;
        if (j < F) {
            o[j] = x_s[s * F + j];
        } else {
            const int c = j - F;
            const float* mo = moments + s * 5 * C + c;
            const MomentStats st = wide_moment_stats(mo[0], mo[C], mo[3 * C], mo[4 * C]);
            const float v[4] = {st.mean, st.std, st.skew, st.kurt};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bf16 hi = __float2bfloat16_rn(v[q]);
                o[F + q * C + c] = hi;
                if (with_lo) {
                    const float rem = v[q] - __bfloat162float(hi);       // inf - inf never occurs: nan_to_num clamps to FLT_MAX
                    o[K9 + q * C + c] = __float2bfloat16_rn(rem == rem ? rem : 0.f);
                }
            }
        }
    }
}
// out[r] = [hi(x[r]) | lo(x[r])] (bf16 [R, 2C], row stride ldo): x = hi + lo to ~2^-17, so a bf16 GEMM over the
// doubled contraction [hi | lo] . [W | W]^T reproduces the fp32 operand (node-level operands only: O(S + T) rows)
__global__ void k_wide_split(const float* __restrict__ x, int ldx, long long R, int C, bf16* __restrict__ out, int ldo) {
    const int half = C >> 1;
    const long long total = R * half;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / half;
        const int c = 2 * (int)(i - r * half);
        const float2 v = *reinterpret_cast<const float2*>(x + r * ldx + c);
        const bf162 hi = __floats2bfloat162_rn(v.x, v.y);
        const float2 hf = __bfloat1622float2(hi);
        *reinterpret_cast<bf162*>(out + r * ldo + c) = hi;
        *reinterpret_cast<bf162*>(out + r * ldo + C + c) = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);
    }
}
// moments backward: dh [S, 9F] fp32 (gradient of hcat) -> dx_s [S,F] bf16 and the per-fibre cubic
// coefficients coef [S,4,C]: dm = A0 + A1 m + A2 d^2 + A3 d^3, already divided by the count
__global__ void k_wide_source_coef(const SegDesc sd, const float* __restrict__ dh, const float* __restrict__ moments, int S, int F,
                                   bf16* __restrict__ dx_s, float* __restrict__ coef) {
    const int C = 2 * F, K9 = 9 * F;
    const long long total = (long long)S * (F + C);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / (F + C);
        const int j = (int)(i - s * (F + C));
        const float* d = dh + s * K9;
        if (j < F) {
            dx_s[s * F + j] = __float2bfloat16_rn(d[j]);
            continue;
        }
        const int c = j - F;
        const float* mo = moments + s * 5 * C + c;
        const float mean = mo[0], ex2 = mo[C], c2 = mo[2 * C], c3 = mo[3 * C], c4 = mo[4 * C];
        const MomentStats st = wide_moment_stats(mean, ex2, c3, c4);
        const float cnt = (float)max(seg_len(sd, (int)s), 1);
        // gradients w.r.t. the four statistics; zero where nan_to_num replaced the value
        const float s3 = st.std0 * st.std0 * st.std0, s4 = s3 * st.std0;
        const bool ok_mean = mean == mean && !isinf(mean);
        const float var = st.var_raw > 0.f ? st.var_raw : 0.01f * st.var_raw;
        const bool ok_var = var == var && !isinf(var);
        const float skew_raw = c3 / s3, kurt_raw = c4 / s4;
        const bool ok_skew = skew_raw == skew_raw && !isinf(skew_raw);
        const bool ok_kurt = kurt_raw == kurt_raw && !isinf(kurt_raw);
        const float d_mean = ok_mean ? d[F + c] : 0.f;
        const float d_std = d[F + C + c];
        const float d_skew = ok_skew ? d[F + 2 * C + c] : 0.f;
        const float d_kurt = ok_kurt ? d[F + 3 * C + c] : 0.f;
        const float d_c3 = d_skew / s3, d_c4 = d_kurt / s4;
        // std (output) = sqrt(nan_to_num(var) + eps); skew, kurt use std0 = sqrt(var + eps) (same value when finite)
        float d_var = ok_var ? d_std / (2.f * st.std) : 0.f;
        d_var += (-3.f * c3 / (s4) * d_skew - 4.f * c4 / (s4 * st.std0) * d_kurt) / (2.f * st.std0);
        const float d_var_raw = d_var * (st.var_raw > 0.f ? 1.f : 0.01f);
        const float d_mu = d_mean - 2.f * mean * d_var_raw - 3.f * c2 * d_c3 - 4.f * c3 * d_c4;
        float* o = coef + s * 4 * C + c;
        o[0] = d_mu / cnt;
        o[C] = 2.f * d_var_raw / cnt;
        o[2 * C] = 3.f * d_c3 / cnt;
        o[3 * C] = 4.f * d_c4 / cnt;
    }
}
// dm[e] = A0[src] + A1[src] m + A2[src] d^2 + A3[src] d^3, d = m - mean[src]   (bf16 [E, C])
template <class TM>
__global__ void __launch_bounds__(256) k_wide_source_dm(const TM* __restrict__ m, const float* __restrict__ moments,
                                                        const float* __restrict__ coef, const int* __restrict__ src, int T,
                                                        long long E, int C, bf16* __restrict__ dm) {
    const int groups = C >> 3;
    const long long total = E * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / groups;
        const int c = 8 * (int)(i - e * groups);
        const long long s = src ? src[e] : e / T;
        float mm[8], mean[8], a0[8], a1[8], a2[8], a3[8], o[8];
        ld8_any(m + e * C + c, mm);
        ld8f(moments + s * 5 * C + c, mean);
        const float* cf = coef + s * 4 * C + c;
        ld8f(cf, a0); ld8f(cf + C, a1); ld8f(cf + 2 * C, a2); ld8f(cf + 3 * C, a3);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float d = mm[q] - mean[q];
            o[q] = a0[q] + a1[q] * mm[q] + (a2[q] + a3[q] * d) * d * d;
        }
        st8(dm + e * C + c, o);
    }
}
// the same per fibre: a block owns a fibre, a thread keeps the fibre's five coefficient rows of its 8 columns in
// registers and strides over the fibre's edges (the gathers by src otherwise dominate the L1/L2 traffic)
template <class TM>
__global__ void __launch_bounds__(256) k_wide_source_dm_seg(const SegDesc sd, const TM* __restrict__ m,
                                                            const float* __restrict__ moments, const float* __restrict__ coef,
                                                            int C, bf16* __restrict__ dm) {
    const int seg = blockIdx.x;
    const int len = seg_len(sd, seg);
    const int groups = C >> 3;
    const int lanes = blockDim.x / groups;
    const int lr = threadIdx.x / groups, c = 8 * (threadIdx.x - lr * groups);
    if (lr >= lanes || len == 0) return;
    float mean[8], a0[8], a1[8], a2[8], a3[8];
    ld8f(moments + (size_t)seg * 5 * C + c, mean);
    const float* cf = coef + (size_t)seg * 4 * C + c;
    ld8f(cf, a0); ld8f(cf + C, a1); ld8f(cf + 2 * C, a2); ld8f(cf + 3 * C, a3);
    for (int i = lr; i < len; i += lanes) {
        const long long e = seg_row(sd, seg, i);
        float mm[8], o[8];
        ld8_any(m + e * C + c, mm);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float d = mm[q] - mean[q];
            o[q] = a0[q] + a1[q] * mm[q] + (a2[q] + a3[q] * d) * d * d;
        }
        st8(dm + e * C + c, o);
    }
}
// out[e] = tab[idx[e]] * (act[e] > 0 ? 1 : slope)    (TModel backward: dht = dasum[tgt] . lrelu'(ht))
__global__ void __launch_bounds__(256) k_wide_gather_mask(const float* __restrict__ tab, const int* __restrict__ idx, int mod,
                                                          const bf16* __restrict__ act, long long E, int C,
                                                          bf16* __restrict__ out) {
    const int groups = C >> 3;
    const long long total = E * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / groups;
        const int c = 8 * (int)(i - e * groups);
        const long long r = idx ? idx[e] : e % mod;
        float t[8], a[8], o[8];
        ld8f(tab + r * C + c, t);
        ld8(act + e * C + c, a);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = t[q] * (a[q] > 0.f ? 1.f : 0.1f);
        st8(out + e * C + c, o);
    }
}

// ------------------------------------------------------------------------------------------------
// time head (reference src/gnn.py:307-312): pred = a . w2 + b2 over the hidden activation a [E,F] (bf16),
// time = softplus(pred) * scale; optional integer outputs visits = rint(time / hours[tgt]), time_int = visits * hours
// (DESIGN.md section 8).  A group of F/8 lanes owns a row (16-byte loads), dot product by shuffles.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wide_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch threshold 20
__global__ void __launch_bounds__(256) k_wide_head_fwd(const bf16* __restrict__ a, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float scale, long long E, int F,
                                                       const float* __restrict__ hours, const int* __restrict__ tgt, int T,
                                                       float* __restrict__ pred_out, float* __restrict__ time,
                                                       float* __restrict__ visits, float* __restrict__ time_int) {
    const int gl = F >> 3;                       // lanes per row (power of two <= 32 is required by the host)
    const int rows_per_warp = 32 / gl;
    const int lane = threadIdx.x & 31, sub = lane / gl, lc = lane - sub * gl;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float w[8];
    ld8f(w2 + 8 * lc, w);
    const float bias = b2[0];
    for (long long r0 = warp0 * rows_per_warp; r0 < E; r0 += nwarps * rows_per_warp) {
        const long long e = r0 + sub;
        float s = 0.f;
        if (e < E) {
            float v[8];
            ld8(a + e * F + 8 * lc, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) s = fmaf(v[q], w[q], s);
        }
        for (int o = gl >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (e < E && lc == 0) {
            const float p = s + bias;
            const float t = wide_softplus(p) * scale;
            pred_out[e] = p;
            time[e] = t;
            if (visits) {
                const float h = hours[tgt ? tgt[e] : (int)(e % T)];
                const float vs = rintf(t / h);
                visits[e] = vs;
                time_int[e] = vs * h;
            }
        }
    }
}
// backward: gp[e] = g[e] * sigmoid(pred[e]) * scale; da[e][:] = gp[e] * w2[:] * lrelu'(a[e][:])  (bf16 [E,F])
__global__ void __launch_bounds__(256) k_wide_head_bwd(const bf16* __restrict__ a, const float* __restrict__ w2,
                                                       const float* __restrict__ pred, const float* __restrict__ g, float scale,
                                                       long long E, int F, float* __restrict__ gp, bf16* __restrict__ da) {
    const int groups = F >> 3;
    const long long total = E * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / groups;
        const int c = 8 * (int)(i - e * groups);
        const float p = pred[e];
        const float sg = p > 20.f ? 1.f : 1.f / (1.f + expf(-p));
        const float coeff = g[e] * sg * scale;
        if (c == 0) gp[e] = coeff;
        float v[8], w[8], o[8];
        ld8(a + e * F + c, v);
        ld8f(w2 + c, w);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = coeff * w[q] * (v[q] > 0.f ? 1.f : 0.1f);
        st8(da + e * F + c, o);
    }
}

// ------------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------------
template <class TI, class TO>
__global__ void k_wide_cast(const TI* __restrict__ in, long long n, TO* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (TO)(float)in[i];
}
// out[c][r] = in[r][c]  (bf16, in [R, C] with leading dimension ld)
__global__ void k_wide_transpose(const bf16* __restrict__ in, int R, int C, int ld, bf16* __restrict__ out) {
    __shared__ bf16 tile[32][34];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[y][threadIdx.x] = in[(size_t)r * ld + c];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int c = c0 + y, r = r0 + threadIdx.x;
        if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][y];
    }
}

}  // namespace pfs
