// edge_fwd_tc.cuh -- EdgeModel forward (reference src/gnn.py:98-101) with BOTH layers of the edge MLP on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split), chained on chip.
//
//   z = W2 . lrelu(P_s[src] + P_t[tgt] + W1_e . x_e) + b2            per edge, + BatchNorm statistics of z
//
// A CTA of 256 threads owns a tile of <= 256 edges (whole fibres, canonical dense order); thread t is edge t, warps
// 0-3 / 4-7 are the two 128-row UMMA blocks.  Per tile:
//   1. every thread writes its x_e row (prefetched one tile ahead) as hi / lo into the K-major core-matrix A operand;
//   2. one thread issues GEMM1 [128 x K=F] . [F x 4F] per block (3 products per k-step) into TMEM and commits;
//   3. epilogue 1: each thread reads its 4F accumulator columns (tcgen05.ld), adds the staged node-table rows
//      P_s[src] + P_t[tgt], applies LeakyReLU and writes the hidden activation hi / lo as the A operand of GEMM2
//      (the layer-1 operand is dead by then and is overlaid) -- the activation never leaves the SM;
//   4. GEMM2 [128 x 4F] . [4F x F]; epilogue 2: bias, store z, running BatchNorm statistics.
// The split: hi = rna_tf32(x), lo = rna_tf32(x - hi), x = hi + lo to 2^-22 |x|.  Products a_lo b_hi + a_hi b_lo + a_hi b_hi.
// The FFMA2 kernel (edge_model.cuh: k_edge_fwd) issues ~1200 instructions per edge; this one ~400, none of them
// multiply-accumulates.  General (CSR) edge lists keep the FFMA2 kernel.
#pragma once
#include "edge_model.cuh"
#include "tc_ptx.cuh"

namespace pfs {

// round-to-nearest split: hi = rna_tf32(x), lo = rna_tf32(x - hi).  Clearing the low mantissa bits instead (one LOP3
// less per element, the tensor core truncating lo) leaves a BIASED 2^-20 error per product; measured on the full-size
// C2 graph that pushed grad x_s to 4e-3 of its fp64 value through the ill-conditioned moment statistics downstream.
__device__ __forceinline__ float tf32_hi(float x) { return to_tf32(x); }
__device__ __forceinline__ float tf32_lo(float x, float hi) { return to_tf32(x - hi); }

template <int F>
struct EdgeFwdTc {
    static constexpr int H = 4 * F;
    static constexpr int KP1 = (F + 7) / 8 * 8;            // K of GEMM1 (multiple of the tf32 MMA K)
    static constexpr int NP1 = (H + 15) / 16 * 16;         // N of GEMM1 (M = 128 needs N % 16 == 0)
    static constexpr int NP2 = (F + 15) / 16 * 16;         // N of GEMM2
    static constexpr int KC1 = KP1 / 4, KC2 = H / 4;       // 16-byte K chunks
    static_assert(H % 8 == 0, "K of GEMM2 must be a multiple of 8");
    static constexpr int LBO_A = 128 * 16;                 // 128 rows per chunk
    static constexpr int LBO_B1 = (NP1 / 8) * 128, LBO_B2 = (NP2 / 8) * 128;
    static constexpr int A_BLOCK = KC2 * LBO_A / 4;        // floats of one (block, hi|lo) operand; GEMM1's overlays its start
    static constexpr int B1_FLOATS = KC1 * LBO_B1 / 4, B2_FLOATS = KC2 * LBO_B2 / 4;
    static constexpr int DCOLS = NP1 + NP2;                // TMEM columns per block: D1 then D2
    static constexpr int TMEM_COLS = 2 * DCOLS <= 32 ? 32 : 2 * DCOLS <= 64 ? 64 : 2 * DCOLS <= 128 ? 128 : 2 * DCOLS <= 256 ? 256 : 512;
    static constexpr int PCP = H + 4;                      // padded class-table row (bank stagger)
    __host__ __device__ static constexpr size_t floats(int max_fib, int T) {
        return 4 * (size_t)A_BLOCK + 2 * B1_FLOATS + 2 * B2_FLOATS + (size_t)max_fib * H + (size_t)T * PCP +
               kWarps * (2 * F + 1) + 16;
    }
    static constexpr bool fits = 2 * DCOLS <= 512 && KC1 <= KC2;
};

template <int F>
__global__ void __launch_bounds__(kThreads, 2) k_edge_fwd_tc(const EdgeFwdParams p) {
    using TC = EdgeFwdTc<F>;
    constexpr int H = TC::H, KC1 = TC::KC1, KC2 = TC::KC2, NP1 = TC::NP1, NP2 = TC::NP2, PCP = TC::PCP;
    extern __shared__ __align__(128) float sm[];
    float* A = sm;                                   // [block 0 hi | block 0 lo | block 1 hi | block 1 lo], A_BLOCK each
    float* B1hi = A + 4 * TC::A_BLOCK;
    float* B1lo = B1hi + TC::B1_FLOATS;
    float* B2hi = B1lo + TC::B1_FLOATS;
    float* B2lo = B2hi + TC::B2_FLOATS;
    float* PS = B2lo + TC::B2_FLOATS;                // [max_fib][H]
    float* PT = PS + (size_t)p.max_fib * H;          // [T][PCP]
    float* red = PT + (size_t)p.tp.T * PCP;          // statistics scratch
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + kWarps * (2 * F + 1) + ((kWarps * (2 * F + 1)) & 1));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const Topo& tp = p.tp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int blk = tid >> 7, r = tid & 127;

    if (tid == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, TC::TMEM_COLS);
    // B operands (weights), hi / lo, K-major core-matrix layout: element (row n, chunk kc) at kc * LBO + n * 16
    for (int i = tid; i < NP1 * KC1; i += kThreads) {
        const int j = i / KC1, kc = i - j * KC1;
        float hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = 4 * kc + q;
            const float w = (j < H && k < F) ? __ldg(p.w1 + (size_t)j * H + 2 * F + k) : 0.f;
            hi[q] = tf32_hi(w);
            lo[q] = tf32_lo(w, hi[q]);
        }
        const int o = (kc * TC::LBO_B1 + j * 16) >> 2;
        *reinterpret_cast<float4*>(B1hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(B1lo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int i = tid; i < NP2 * KC2; i += kThreads) {
        const int f = i / KC2, kc = i - f * KC2;
        float hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float w = f < F ? __ldg(p.w2 + (size_t)f * H + 4 * kc + q) : 0.f;
            hi[q] = tf32_hi(w);
            lo[q] = tf32_lo(w, hi[q]);
        }
        const int o = (kc * TC::LBO_B2 + f * 16) >> 2;
        *reinterpret_cast<float4*>(B2hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(B2lo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    float b2r[F];
#pragma unroll
    for (int j = 0; j < F; ++j) b2r[j] = __ldg(p.b2 + j);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc1 = umma_idesc_tf32(128, NP1), idesc2 = umma_idesc_tf32(128, NP2);
    float* Ahi = A + (size_t)blk * 2 * TC::A_BLOCK;
    float* Alo = Ahi + TC::A_BLOCK;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(blk * TC::DCOLS);

    const int total = tp.ntiles * tp.G;
    const int t_begin = chunk_begin(blockIdx.x, gridDim.x, total), t_end = chunk_begin(blockIdx.x + 1, gridDim.x, total);
    RunningStats<F> rs;
    rs.reset();
    float* rec = p.bn_partial ? p.bn_partial + (size_t)blockIdx.x * p.nrec * bn_partial_stride(F) : nullptr;
    int slot = 0, cur_graph = -1;
    uint32_t parity = 0;

    // inputs of the next tile, fetched one tile ahead into registers: the thread's x_e row and its share of the
    // tile's node-table rows (P_s rows of the tile's fibres, the graph's P_t)
    constexpr int H4 = H / 4;
    float xn[F];
    float4 psn[2], ptn;
    auto prefetch = [&](int tile) {
        if (tile >= t_end) return;
        const Tile t = get_tile(tp, tile);
        if (tid < t.ne) load_row<F>(p.x_e + ((size_t)t.g * tp.E + t.q0 + tid) * F, xn);
        const float4* ps = reinterpret_cast<const float4*>(p.Ps + ((size_t)t.g * tp.S + t.fibre0) * H);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = tid + q * kThreads;
            if (i < t.nfib * H4) psn[q] = __ldg(ps + i);
        }
        if (tid < tp.T * H4) ptn = __ldg(reinterpret_cast<const float4*>(p.Pt + (size_t)t.g * tp.T * H) + tid);
    };
    prefetch(t_begin);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const Tile t = get_tile(tp, tile);
        const bool active = tid < t.ne;
        if (rec && t.g != cur_graph) {   // graph boundary: emit the finished graph's statistics
            if (cur_graph >= 0) rs.flush(red, rec + (size_t)(slot++) * bn_partial_stride(F), cur_graph);
            cur_graph = t.g;
        }
        // ---- layer-1 A operand (x_e hi / lo) and the staged node tables ------------------------------------
        if (active) {
#pragma unroll
            for (int kc = 0; kc < KC1; ++kc) {
                float hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = 4 * kc + q;
                    const float v = k < F ? xn[k < F ? k : 0] : 0.f;
                    hi[q] = tf32_hi(v);
                    lo[q] = tf32_lo(v, hi[q]);
                }
                const int o = (kc * TC::LBO_A + r * 16) >> 2;
                *reinterpret_cast<float4*>(Ahi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(Alo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = tid + q * kThreads;
            if (i < t.nfib * H4) reinterpret_cast<float4*>(PS)[i] = psn[q];
        }
        if (tid < tp.T * H4) {
            const int c = tid / H4, k4 = tid - c * H4;
            *reinterpret_cast<float4*>(PT + c * PCP + 4 * k4) = ptn;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                if (b * 128 >= t.ne) break;
                const uint32_t ah = smem_u32(A + (size_t)b * 2 * TC::A_BLOCK), al = ah + TC::A_BLOCK * 4;
                const uint32_t bh = smem_u32(B1hi), bl = smem_u32(B1lo);
#pragma unroll
                for (int ks = 0; ks < TC::KP1 / 8; ++ks) {
                    const uint32_t ao = ks * 2 * TC::LBO_A, bo = ks * 2 * TC::LBO_B1;
                    const uint64_t dah = umma_desc(ah + ao, TC::LBO_A), dal = umma_desc(al + ao, TC::LBO_A);
                    const uint64_t dbh = umma_desc(bh + bo, TC::LBO_B1), dbl = umma_desc(bl + bo, TC::LBO_B1);
                    umma_tf32(tmem + b * TC::DCOLS, dal, dbh, idesc1, ks > 0 ? 1u : 0u);   // small terms first
                    umma_tf32(tmem + b * TC::DCOLS, dah, dbl, idesc1, 1u);
                    umma_tf32(tmem + b * TC::DCOLS, dah, dbh, idesc1, 1u);
                }
            }
            umma_commit(bars);
        }
        prefetch(tile + 1);               // in flight under the MMAs and both epilogues
        // ---- epilogue 1: h = acc + P_s[src] + P_t[tgt], a1 = lrelu(h) -> layer-2 A operand -------------------
        mbar_wait(bars, parity);
        tc_fence_after();
        const bool warp_live = warp * 32 < t.ne;    // tcgen05.ld is warp-collective: the whole warp takes part or none of it
        if (warp_live) {
            const int lf = active ? tid / tp.T : 0, tg = active ? tid - lf * tp.T : 0;
            const float* ps = PS + lf * H;
            const float* pt = PT + tg * PCP;
#pragma unroll
            for (int c0 = 0; c0 < H; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                if (!active) continue;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    if (c0 + 4 * q4 >= H) break;
                    const float4 a = *reinterpret_cast<const float4*>(ps + c0 + 4 * q4);
                    const float4 b = *reinterpret_cast<const float4*>(pt + c0 + 4 * q4);
                    const float h0 = lrelu(v[4 * q4] + a.x + b.x), h1 = lrelu(v[4 * q4 + 1] + a.y + b.y);
                    const float h2 = lrelu(v[4 * q4 + 2] + a.z + b.z), h3 = lrelu(v[4 * q4 + 3] + a.w + b.w);
                    const float g0 = tf32_hi(h0), g1 = tf32_hi(h1), g2 = tf32_hi(h2), g3 = tf32_hi(h3);
                    const int o = (((c0 >> 2) + q4) * TC::LBO_A + r * 16) >> 2;
                    if (p.act_save)
                        *reinterpret_cast<float4*>(p.act_save + ((size_t)t.g * tp.E + t.q0 + tid) * H + c0 + 4 * q4) = make_float4(h0, h1, h2, h3);
                    *reinterpret_cast<float4*>(Ahi + o) = make_float4(g0, g1, g2, g3);
                    *reinterpret_cast<float4*>(Alo + o) = make_float4(tf32_lo(h0, g0), tf32_lo(h1, g1), tf32_lo(h2, g2), tf32_lo(h3, g3));
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                if (b * 128 >= t.ne) break;
                const uint32_t ah = smem_u32(A + (size_t)b * 2 * TC::A_BLOCK), al = ah + TC::A_BLOCK * 4;
                const uint32_t bh = smem_u32(B2hi), bl = smem_u32(B2lo);
#pragma unroll
                for (int ks = 0; ks < H / 8; ++ks) {
                    const uint32_t ao = ks * 2 * TC::LBO_A, bo = ks * 2 * TC::LBO_B2;
                    const uint64_t dah = umma_desc(ah + ao, TC::LBO_A), dal = umma_desc(al + ao, TC::LBO_A);
                    const uint64_t dbh = umma_desc(bh + bo, TC::LBO_B2), dbl = umma_desc(bl + bo, TC::LBO_B2);
                    umma_tf32(tmem + b * TC::DCOLS + NP1, dal, dbh, idesc2, ks > 0 ? 1u : 0u);
                    umma_tf32(tmem + b * TC::DCOLS + NP1, dah, dbl, idesc2, 1u);
                    umma_tf32(tmem + b * TC::DCOLS + NP1, dah, dbh, idesc2, 1u);
                }
            }
            umma_commit(bars + 1);
        }
        // ---- epilogue 2: z = acc + b2, store, statistics -------------------------------------------------------
        mbar_wait(bars + 1, parity);
        parity ^= 1;
        tc_fence_after();
        if (warp_live) {
            float z[F];
#pragma unroll
            for (int c0 = 0; c0 < F; c0 += 16) {
                float v[16];
                tmem_ld16(trow + NP1 + c0, v);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (c0 + q < F) z[c0 + q] = v[q] + b2r[c0 + q];
            }
            if (active) {
                store_row<F>(p.z_out + ((size_t)t.g * tp.E + t.q0 + tid) * F, z);
                if (rec) rs.add(z);
            }
        }
        tc_fence_before();      // the next tile's MMAs overwrite the accumulators every thread has just read
    }
    if (rec) {
        if (cur_graph >= 0) rs.flush(red, rec + (size_t)(slot++) * bn_partial_stride(F), cur_graph);
        for (; slot < p.nrec; ++slot)      // unused records: count 0
            if (tid == 0) rec[(size_t)slot * bn_partial_stride(F) + 2 * F] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TC::TMEM_COLS);
}

}  // namespace pfs
