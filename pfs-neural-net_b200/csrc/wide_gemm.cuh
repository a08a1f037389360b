// wide_gemm.cuh -- bf16 tensor-core GEMMs of the wide-feature path (Fdim >= 32, BASELINE configs C4/C5b).
//
// At Fdim 128 every MLP layer of the message-passing block (reference src/gnn.py:65-71 and its users
// :100,:136,:153,:188) is a real dense contraction, so it runs on tcgen05 with TMEM accumulators:
//
//   k_wide_gemm_nt : C[M,N] = epilogue(A[M,K] . B[N,K]^T)      -- forward layers and input gradients
//       A = per-edge / per-node rows (row-major, K contiguous), B = torch Linear weight [out,in].
//       Fused epilogue: + bias (optionally row-scaled) + gathered node tables (P_s[src] + P_t[tgt],
//       the first-layer split of DESIGN.md 3.1), LeakyReLU, multiplication by the LeakyReLU
//       derivative recovered from the saved activation, bf16 (TMA store) and/or fp32 output.
//   k_wide_gemm_tn : W[J,Kx] = D[E,J]^T . X[E,Kx]              -- weight gradients (contraction over edges)
//       both operands are read in their natural row-major layout as MN-major UMMA operands; the
//       edge range is split over CTAs, partials are summed in a fixed order (deterministic).
//
// Structure (both): warp 0 = TMA producer (128-byte-swizzled boxes into a STAGES-deep ring),
// warp 1 = single-thread tcgen05.mma issuer, warps 2-5 = epilogue (tcgen05.ld of their 32 TMEM
// lanes).  NT is persistent over output tiles with a double-buffered accumulator so the epilogue
// of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "tc_ptx.cuh"

namespace pfs {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;        // 64 bf16 = 128 bytes = one swizzle row
constexpr int kGemmThreads = 192;
constexpr float kWideSlope = 0.1f;

struct GemmEpilogue {
    const float* bias;             // [N] or null
    const float* bias_rowscale;    // [M] or null: bias[n] * bias_rowscale[m]
    const float* tab0;             // [R0, N] fp32 row table gathered by idx0[m] (or m / div0), or null
    const int* idx0;
    int div0;
    const float* tab1;             // [R1, N] fp32 row table gathered by idx1[m] (or m % mod1), or null
    const int* idx1;
    int mod1;
    const __nv_bfloat16* mask;     // [M, ldmask] saved activation: multiply by (mask > 0 ? 1 : slope)
    int ldmask;
    int act;                       // 1: LeakyReLU(0.1)
    float* out_f32;                // optional fp32 output [M, ldf]
    int ldf;
    int out_bf16;                  // 1: bf16 output through the TMA store map
};

template <int BN, int STAGES>
struct GemmNtSmem {
    static constexpr int kA = kGemmBM * kGemmBK * 2;     // 16 KB
    static constexpr int kB = BN * kGemmBK * 2;
    static constexpr int kC = kGemmBM * 64 * 2;          // staging of a 128 x 64 bf16 output box
    static constexpr int kBars = 256;
    static constexpr size_t bytes = (size_t)STAGES * (kA + kB) + 2 * kC + kBars + 1024;   // + alignment slack
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_wide_gemm_nt(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmEpilogue ep, int M, int N, int K) {
    using SM = GemmNtSmem<BN, STAGES>;
    extern __shared__ uint8_t gemm_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + STAGES * SM::kA;
    uint8_t* sC = sB + STAGES * SM::kB;
    uint64_t* full = reinterpret_cast<uint64_t*>(sC + 2 * SM::kC);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = (M + kGemmBM - 1) / kGemmBM, nt = (N + BN - 1) / BN;
    const int tiles = mt * nt;
    const int KB = (K + kGemmBK - 1) / kGemmBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull + s, 1);
            mbar_init(tempty + s, 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (t / nt) * kGemmBM, n0 = (t % nt) * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty + stage, ph ^ 1);
                    mbar_expect_tx(full + stage, SM::kA + SM::kB);
                    tma_load_2d(sA + stage * SM::kA, &tmA, full + stage, kb * kGemmBK, m0);
                    tma_load_2d(sB + stage * SM::kB, &tmB, full + stage, kb * kGemmBK, n0);
                    if (++stage == STAGES) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kGemmBM, BN, 0, 0);
            int stage = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(tempty + as, aph ^ 1);        // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem + as * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(full + stage, ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * SM::kA), b0 = smem_u32(sB + stage * SM::kB);
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; ++k) {
                        const uint64_t da = umma_desc_sw128(a0 + k * 32, 16, 1024);
                        const uint64_t db = umma_desc_sw128(b0 + k * 32, 16, 1024);
                        umma_f16(d, da, db, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(empty + stage);          // frees the smem slot once the MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
                umma_commit(tfull + as);                 // accumulator complete
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lanes 32 * (warp % 4) .. + 31 =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const bool issuer = threadIdx.x == 64;          // first epilogue thread issues the TMA stores
        int it = 0, chunk = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int m0 = (t / nt) * kGemmBM, n0 = (t % nt) * BN;
            const int m = m0 + row;
            const bool row_ok = m < M;
            mbar_wait(tfull + as, aph);
            tc_fence_after();
            const float* t0 = nullptr;
            const float* t1 = nullptr;
            float rs = 1.f;
            if (row_ok) {
                if (ep.tab0) t0 = ep.tab0 + (size_t)(ep.idx0 ? ep.idx0[m] : m / ep.div0) * N;
                if (ep.tab1) t1 = ep.tab1 + (size_t)(ep.idx1 ? ep.idx1[m] : m % ep.mod1) * N;
                if (ep.bias_rowscale) rs = ep.bias_rowscale[m];
            }
            const int ncols = min(BN, N - n0);
            for (int c64 = 0; c64 < ncols; c64 += 64, ++chunk) {
                uint8_t* stg = sC + (chunk & 1) * SM::kC;
                if (ep.out_bf16) {
                    if (issuer) tma_store_wait_read<1>();    // the store that last read this buffer is done
                    named_bar_sync(1, 128);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c0 = c64 + 32 * h;
                    if (c0 >= ncols) break;
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + as * BN + c0, v);
                    const int n = n0 + c0;
                    if (row_ok) {
#pragma unroll
                        for (int q = 0; q < 32; q += 4) {
                            if (n + q + 4 <= N) {
                                if (ep.bias) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n + q));
                                    v[q] = fmaf(b.x, rs, v[q]); v[q + 1] = fmaf(b.y, rs, v[q + 1]);
                                    v[q + 2] = fmaf(b.z, rs, v[q + 2]); v[q + 3] = fmaf(b.w, rs, v[q + 3]);
                                }
                                if (t0) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(t0 + n + q));
                                    v[q] += b.x; v[q + 1] += b.y; v[q + 2] += b.z; v[q + 3] += b.w;
                                }
                                if (t1) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(t1 + n + q));
                                    v[q] += b.x; v[q + 1] += b.y; v[q + 2] += b.z; v[q + 3] += b.w;
                                }
                            }
                        }
                        if (ep.act) {
#pragma unroll
                            for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], kWideSlope * v[q]);
                        }
                        if (ep.mask) {
                            const __nv_bfloat16* mp = ep.mask + (size_t)m * ep.ldmask + n;
#pragma unroll
                            for (int q = 0; q < 32; q += 8) {
                                if (n + q + 8 <= N) {
                                    const uint4 w = __ldg(reinterpret_cast<const uint4*>(mp + q));
                                    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        // bf16 sign/zero test on the raw bits: value > 0 <=> not negative and not zero
                                        const uint32_t lo = ww[e] & 0xFFFFu, hi = ww[e] >> 16;
                                        const bool plo = (lo & 0x8000u) == 0 && (lo & 0x7FFFu) != 0;
                                        const bool phi = (hi & 0x8000u) == 0 && (hi & 0x7FFFu) != 0;
                                        v[q + 2 * e] *= plo ? 1.f : kWideSlope;
                                        v[q + 2 * e + 1] *= phi ? 1.f : kWideSlope;
                                    }
                                }
                            }
                        }
                        if (ep.out_f32) {
                            float* op = ep.out_f32 + (size_t)m * ep.ldf + n;
#pragma unroll
                            for (int q = 0; q < 32; q += 4)
                                if (n + q + 4 <= N)
                                    *reinterpret_cast<float4*>(op + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                        }
                    }
                    if (ep.out_bf16) {
                        // staging row = 128 bytes (64 bf16), 16-byte chunk c stored at c ^ (row % 8): the
                        // 128-byte swizzle the TMA store map expects, and bank-conflict free
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint4 w = make_uint4(pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                                                       pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
                            const int cc = (4 * h + c) ^ (row & 7);
                            *reinterpret_cast<uint4*>(stg + row * 128 + cc * 16) = w;
                        }
                    }
                }
                if (ep.out_bf16) {
                    fence_proxy_async();
                    named_bar_sync(1, 128);
                    if (issuer) {
                        tma_store_2d(&tmC, stg, n0 + c64, m0);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + as);
        }
        if (issuer) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 2 * BN);
}

// -------------------------------------------------------------------------------------------------
// W_partial[split][J][Kx] = sum over the split's rows e of D[e][j] * X[e][k]
// grid = (J tiles of 128) * (Kx tiles of BN), splits
// -------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct GemmTnSmem {
    static constexpr int kA = kGemmBK * kGemmBM * 2;     // 64 rows x 128 features: two 8 KB boxes
    static constexpr int kB = kGemmBK * BN * 2;          // BN / 64 boxes of 8 KB
    static constexpr size_t bytes = (size_t)STAGES * (kA + kB) + 256 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_wide_gemm_tn(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX,
               float* __restrict__ partial, int E, int J, int Kx, int rows_per_split) {
    using SM = GemmTnSmem<BN, STAGES>;
    extern __shared__ uint8_t gemm_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + STAGES * SM::kA;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * SM::kB);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = (Kx + BN - 1) / BN;
    const int j0 = (blockIdx.x / kt) * kGemmBM, k0 = (blockIdx.x % kt) * BN;
    const int split = blockIdx.y;
    const int e_begin = split * rows_per_split, e_end = min(E, e_begin + rows_per_split);
    const int KB = e_end > e_begin ? (e_end - e_begin + kGemmBK - 1) / kGemmBK : 0;
    constexpr int kTmemCols = BN < 32 ? 32 : BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmD);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < KB; ++kb) {
                const int e0 = e_begin + kb * kGemmBK;
                mbar_wait(empty + stage, ph ^ 1);
                mbar_expect_tx(full + stage, SM::kA + SM::kB);
                // rows beyond e_end belong to the next split: the boxes are clipped to E by TMA (zero
                // fill) and to the split by rows_per_split being a multiple of the box height
#pragma unroll
                for (int b = 0; b < kGemmBM / 64; ++b)
                    tma_load_2d(sA + stage * SM::kA + b * 8192, &tmD, full + stage, j0 + 64 * b, e0);
#pragma unroll
                for (int b = 0; b < BN / 64; ++b)
                    tma_load_2d(sB + stage * SM::kB + b * 8192, &tmX, full + stage, k0 + 64 * b, e0);
                if (++stage == STAGES) {
                    stage = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kGemmBM, BN, 1, 1);
            int stage = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(full + stage, ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + stage * SM::kA), b0 = smem_u32(sB + stage * SM::kB);
#pragma unroll
                for (int k = 0; k < kGemmBK / 16; ++k) {
                    // 16 contraction rows = two 8-row groups of 1024 bytes
                    const uint64_t da = umma_desc_sw128(a0 + k * 2048, 8192, 1024);
                    const uint64_t db = umma_desc_sw128(b0 + k * 2048, 8192, 1024);
                    umma_f16(tmem, da, db, idesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(empty + stage);
                if (++stage == STAGES) {
                    stage = 0;
                    ph ^= 1;
                }
            }
            umma_commit(tfull);
        }
    } else {
        const int quarter = warp & 3;
        const int j = j0 + quarter * 32 + lane;
        if (KB > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        float* out = partial + ((size_t)split * J + j) * Kx;
        const int ncols = min(BN, Kx - k0);
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            float v[32];
            if (KB > 0) {
                tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = 0.f;
            }
            if (j < J) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    if (k0 + c0 + q + 4 <= Kx)
                        *reinterpret_cast<float4*>(out + k0 + c0 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// out[j * ldo + coff + k] (+)= sum_s partial[s][j][k], fixed order
__global__ void k_wide_reduce_splits(const float* __restrict__ partial, int splits, int J, int Kx,
                                     float* __restrict__ out, int ldo, int coff, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J * Kx) return;
    float s = 0.f;
    for (int q = 0; q < splits; ++q) s += partial[(size_t)q * J * Kx + i];
    const int j = i / Kx, k = i - j * Kx;
    float* o = out + (size_t)j * ldo + coff + k;
    *o = accumulate ? *o + s : s;
}

}  // namespace pfs
