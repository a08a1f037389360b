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
// warp 1 = single-thread tcgen05.mma issuer, remaining warps = epilogue (tcgen05.ld of their 32 TMEM
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
    int rows0;                     // rows of tab0 (dense addressing is clamped to it)
    const float* tab1;             // [R1, N] fp32 row table gathered by idx1[m] (or m % mod1), or null
    const int* idx1;
    int mod1;
    const __nv_bfloat16* mask;     // [M, ldmask] saved activation: multiply by (mask > 0 ? 1 : slope)
    int ldmask;
    const float* brow;             // [R, N] fp32 rows added per TILE: row = m0 / brow_div (tiles never straddle a row), or null
    int brow_div;
    int k1;                        // columns of the second A operand (tmA2), contracted FIRST against B[:, 0:k1); 0 = none
    int a2_mod;                    // its row for tile m0 is m0 % a2_mod (dense layout: x_t[tgt], tgt = e % T)
    int act;                       // 1: LeakyReLU(0.1)
    float* out_f32;                // optional fp32 output [M, ldf]
    int ldf;
    int out_bf16;                  // 1: bf16 output through the TMA store map
};

constexpr int kGemmEpiWarps = 16;
constexpr int kGemmNtThreads = 64 + 32 * kGemmEpiWarps;      // warp 0 = TMA, warp 1 = MMA, warps 2..17 = epilogue
constexpr int kGemmEpiChunk = 128;       // accumulator columns staged per epilogue round

// B-stationary variants keep the CTA's whole weight tile [BN, K] in shared memory as BT k-blocks (BT = 0: B streams
// through the ring with A)
template <int BN, int STAGES, int BT = 0>
struct GemmNtSmem {
    static constexpr int kA = kGemmBM * kGemmBK * 2;     // 16 KB
    static constexpr int kB = BN * kGemmBK * 2;
    static constexpr int kBT = BT ? BT : STAGES;              // B tiles held in shared memory
    static constexpr int kC = kGemmBM * kGemmEpiChunk * 4;   // fp32 staging of a 128 x 128 accumulator block
    static constexpr int kBars = 256;
    static constexpr size_t bytes = (size_t)STAGES * kA + (size_t)kBT * kB + kC + kBars;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ bool bf16_bits_positive(uint32_t h) { return (h & 0x8000u) == 0 && (h & 0x7FFFu) != 0; }

// Epilogue in two phases per 128-column block, so that TMEM is read in its natural mapping
// (thread = accumulator row) while every global access is coalesced (warp = one row, lane = 4 columns):
//   phase 1: tcgen05.ld -> fp32 staging tile in shared memory (16-byte chunks XOR-swizzled by row)
//   phase 2: staging -> registers, + bias / gathered table rows, LeakyReLU, derivative mask, bf16 / fp32 stores
// BT > 0: every tile of a CTA has the same n block (gridDim.x % nt == 0) and K_total <= 64 BT, so the weight tile is loaded
// once per CTA and only A streams through the ring -- the L2 -> SM traffic per output tile halves for the K <= 256 layers.
template <int BN, int STAGES, bool TABLES, bool MASK, int BT = 0>
__global__ void __launch_bounds__(kGemmNtThreads, 1)
k_wide_gemm_nt(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const GemmEpilogue ep,
               __nv_bfloat16* __restrict__ out_bf16, int ldc, int M, int N, int K) {
    constexpr bool BSTAT = BT > 0;
    using SM = GemmNtSmem<BN, STAGES, BT>;
    extern __shared__ __align__(1024) uint8_t gemm_smem[];   // no static shared memory: the window starts 1024-aligned
    uint8_t* sA = gemm_smem;
    uint8_t* sB = sA + STAGES * SM::kA;
    float* stg = reinterpret_cast<float*>(sB + SM::kBT * SM::kB);
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg) + SM::kC);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* bfull = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = (M + kGemmBM - 1) / kGemmBM, nt = (N + BN - 1) / BN;
    const int tiles = mt * nt;
    // K-concatenation: [A2 | A] . B^T with A2 [*, k1] (rows m0 % a2_mod) and A [M, K]; B is [N, k1 + K]
    const int KB1 = (ep.k1 + kGemmBK - 1) / kGemmBK;
    const int KB = KB1 + (K + kGemmBK - 1) / kGemmBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull + s, 1);
            mbar_init(tempty + s, kGemmEpiWarps);
        }
        mbar_init(bfull, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
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
            if constexpr (BSTAT) {      // the CTA's weight tile, once
                const int n0 = (blockIdx.x % nt) * BN;
                mbar_expect_tx(bfull, KB * SM::kB);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(sB + kb * SM::kB, &tmB, bfull, kb < KB1 ? kb * kGemmBK : ep.k1 + (kb - KB1) * kGemmBK, n0);
            }
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (t / nt) * kGemmBM, n0 = (t % nt) * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty + stage, ph ^ 1);
                    mbar_expect_tx(full + stage, BSTAT ? SM::kA : SM::kA + SM::kB);
                    if (kb < KB1) {     // columns beyond k1 are zero-filled in A2, so the extra B columns contribute nothing
                        tma_load_2d(sA + stage * SM::kA, &tmA2, full + stage, kb * kGemmBK, m0 % ep.a2_mod);
                        if constexpr (!BSTAT) tma_load_2d(sB + stage * SM::kB, &tmB, full + stage, kb * kGemmBK, n0);
                    } else {
                        tma_load_2d(sA + stage * SM::kA, &tmA, full + stage, (kb - KB1) * kGemmBK, m0);
                        if constexpr (!BSTAT)
                            tma_load_2d(sB + stage * SM::kB, &tmB, full + stage, ep.k1 + (kb - KB1) * kGemmBK, n0);
                    }
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
            if constexpr (BSTAT) mbar_wait(bfull, 0);
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(tempty + as, aph ^ 1);        // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem + as * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(full + stage, ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * SM::kA), b0 = smem_u32(sB + (BSTAT ? kb : stage) * SM::kB);
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
        // ===== epilogue warps 2..17 =====
        const int ew = warp - 2;
        const int quarter = warp & 3;                    // TMEM lanes this warp may read: 32 * (warp % 4) ..
        const int cq = ew >> 2;                          // which 32 of the block's 128 columns it drains
        const int trow = quarter * 32 + lane;            // accumulator row of this thread in phase 1
        constexpr int kRows = kGemmBM / kGemmEpiWarps;   // rows per warp in phase 2: ew, ew + 16, ...
        constexpr int kChunks = (BN + kGemmEpiChunk - 1) / kGemmEpiChunk;
        // staging addresses (16-byte chunks XOR-swizzled by the row; rows ew + 16 rr share ew's low bits)
        float4* const p1row = reinterpret_cast<float4*>(stg + trow * kGemmEpiChunk) + 8 * cq;
        const int t7 = trow & 7;
        const float4* const p2row = reinterpret_cast<const float4*>(stg + ew * kGemmEpiChunk) + ((lane & ~7) | ((lane ^ ew) & 7));
        const size_t ostride = (size_t)kGemmEpiWarps * ldc, fstride = (size_t)kGemmEpiWarps * ep.ldf;
        const size_t mstride = (size_t)kGemmEpiWarps * ep.ldmask;
        int it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int m0 = (t / nt) * kGemmBM, n0 = (t % nt) * BN;
            const int ncols = min(BN, N - n0);
            const int mbase = m0 + ew;
            const bool full = m0 + kGemmBM <= M;          // no row of this tile is out of range
            // per-row gather offsets, once per tile (rows beyond M are clamped for the loads, masked at the stores)
            uint32_t off0[kRows], off1[kRows];
            if constexpr (TABLES) {
                if (ep.idx0) {
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr)
                        off0[rr] = (uint32_t)__ldg(ep.idx0 + min(mbase + kGemmEpiWarps * rr, M - 1)) * (uint32_t)N;
                } else {
                    int q = mbase / ep.div0, rem = mbase - q * ep.div0;
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr) {
                        off0[rr] = (uint32_t)min(q, ep.rows0 - 1) * (uint32_t)N;
                        rem += kGemmEpiWarps;
                        while (rem >= ep.div0) { rem -= ep.div0; ++q; }
                    }
                }
                if (ep.idx1) {
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr)
                        off1[rr] = (uint32_t)__ldg(ep.idx1 + min(mbase + kGemmEpiWarps * rr, M - 1)) * (uint32_t)N;
                } else {
                    int rem = mbase % ep.mod1;
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr) {
                        off1[rr] = (uint32_t)rem * (uint32_t)N;
                        rem += kGemmEpiWarps;
                        while (rem >= ep.mod1) rem -= ep.mod1;
                    }
                }
            }
            // the derivative mask does not depend on the accumulator: its loads for a block of columns are issued one
            // block ahead (for the first block: before waiting for the MMAs), so their latency hides behind the tile
            uint2 mkv[kRows];
            auto load_mask = [&](int c0m) {
                const int colm = c0m + 4 * lane;
                if (colm < ncols && n0 + colm + 4 <= N) {
                    const __nv_bfloat16* mp = ep.mask + (size_t)mbase * ep.ldmask + n0 + colm;
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr) {
                        mkv[rr] = make_uint2(0u, 0u);
                        if (full || mbase + kGemmEpiWarps * rr < M) mkv[rr] = __ldg(reinterpret_cast<const uint2*>(mp + rr * mstride));
                    }
                }
            };
            if constexpr (MASK) load_mask(0);
            mbar_wait(tfull + as, aph);
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                const int c0 = ch * kGemmEpiChunk;
                if (c0 >= ncols) break;
                named_bar_sync(1, 32 * kGemmEpiWarps);   // phase 2 of the previous block has left the staging tile
                // ---- phase 1: TMEM -> staging ----
                if (c0 + 32 * cq < ncols) {
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + as * BN + c0 + 32 * cq, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) p1row[q ^ t7] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
                if (c0 + kGemmEpiChunk >= ncols) {       // accumulator fully drained: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + as);
                }
                named_bar_sync(1, 32 * kGemmEpiWarps);
                // ---- phase 2: warp = row, lane = 4 consecutive columns; straight-line, loads first ----
                const int col = c0 + 4 * lane;
                const int n = n0 + col;
                if (col < ncols && n + 4 <= N) {
                    float4 acc[kRows];
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr) acc[rr] = p2row[rr * (kGemmEpiWarps * kGemmEpiChunk / 4)];
                    if (ep.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
                        if (ep.bias_rowscale) {
#pragma unroll
                            for (int rr = 0; rr < kRows; ++rr) {
                                const float rs = __ldg(ep.bias_rowscale + min(mbase + kGemmEpiWarps * rr, M - 1));
                                acc[rr].x = fmaf(b.x, rs, acc[rr].x); acc[rr].y = fmaf(b.y, rs, acc[rr].y);
                                acc[rr].z = fmaf(b.z, rs, acc[rr].z); acc[rr].w = fmaf(b.w, rs, acc[rr].w);
                            }
                        } else {
#pragma unroll
                            for (int rr = 0; rr < kRows; ++rr) {
                                acc[rr].x += b.x; acc[rr].y += b.y; acc[rr].z += b.z; acc[rr].w += b.w;
                            }
                        }
                    }
                    if (ep.brow) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.brow + (size_t)(m0 / ep.brow_div) * N + n));
#pragma unroll
                        for (int rr = 0; rr < kRows; ++rr) {
                            acc[rr].x += b.x; acc[rr].y += b.y; acc[rr].z += b.z; acc[rr].w += b.w;
                        }
                    }
                    if constexpr (TABLES) {
                        // fp32 tables (a bf16 table would flip LeakyReLU signs near zero); gathers in batches of
                        // four rows: eight 16-byte loads in flight per lane
                        const float* t0p = ep.tab0 + n;
                        const float* t1p = ep.tab1 + n;
#pragma unroll
                        for (int r4 = 0; r4 < kRows; r4 += 4) {
                            float4 t0v[4], t1v[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                t0v[q] = __ldg(reinterpret_cast<const float4*>(t0p + off0[r4 + q]));
                                t1v[q] = __ldg(reinterpret_cast<const float4*>(t1p + off1[r4 + q]));
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                acc[r4 + q].x += t0v[q].x + t1v[q].x; acc[r4 + q].y += t0v[q].y + t1v[q].y;
                                acc[r4 + q].z += t0v[q].z + t1v[q].z; acc[r4 + q].w += t0v[q].w + t1v[q].w;
                            }
                        }
                    }
                    if (ep.act) {
#pragma unroll
                        for (int rr = 0; rr < kRows; ++rr) {
                            acc[rr].x = fmaxf(acc[rr].x, kWideSlope * acc[rr].x); acc[rr].y = fmaxf(acc[rr].y, kWideSlope * acc[rr].y);
                            acc[rr].z = fmaxf(acc[rr].z, kWideSlope * acc[rr].z); acc[rr].w = fmaxf(acc[rr].w, kWideSlope * acc[rr].w);
                        }
                    }
                    if constexpr (MASK) {
#pragma unroll
                        for (int rr = 0; rr < kRows; ++rr) {
                            acc[rr].x *= bf16_bits_positive(mkv[rr].x & 0xFFFFu) ? 1.f : kWideSlope;
                            acc[rr].y *= bf16_bits_positive(mkv[rr].x >> 16) ? 1.f : kWideSlope;
                            acc[rr].z *= bf16_bits_positive(mkv[rr].y & 0xFFFFu) ? 1.f : kWideSlope;
                            acc[rr].w *= bf16_bits_positive(mkv[rr].y >> 16) ? 1.f : kWideSlope;
                        }
                        if (ch + 1 < kChunks && c0 + kGemmEpiChunk < ncols) load_mask(c0 + kGemmEpiChunk);   // next block's mask
                    }
                    if (out_bf16) {
                        __nv_bfloat16* op = out_bf16 + (size_t)mbase * ldc + n;
                        if (full) {
#pragma unroll
                            for (int rr = 0; rr < kRows; ++rr)
                                *reinterpret_cast<uint2*>(op + rr * ostride) =
                                    make_uint2(pack_bf16(acc[rr].x, acc[rr].y), pack_bf16(acc[rr].z, acc[rr].w));
                        } else {
#pragma unroll
                            for (int rr = 0; rr < kRows; ++rr)
                                if (mbase + kGemmEpiWarps * rr < M)
                                    *reinterpret_cast<uint2*>(op + rr * ostride) =
                                        make_uint2(pack_bf16(acc[rr].x, acc[rr].y), pack_bf16(acc[rr].z, acc[rr].w));
                        }
                    }
                    if (ep.out_f32) {
                        float* op = ep.out_f32 + (size_t)mbase * ep.ldf + n;
#pragma unroll
                        for (int rr = 0; rr < kRows; ++rr)
                            if (full || mbase + kGemmEpiWarps * rr < M) *reinterpret_cast<float4*>(op + rr * fstride) = acc[rr];
                    }
                }
            }
        }
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
    static constexpr size_t bytes = (size_t)STAGES * (kA + kB) + 256;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_wide_gemm_tn(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX,
               float* __restrict__ partial, int E, int J, int Kx, int rows_per_split) {
    using SM = GemmTnSmem<BN, STAGES>;
    extern __shared__ __align__(1024) uint8_t gemm_smem[];
    uint8_t* sA = gemm_smem;
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
