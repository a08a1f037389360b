// wide.cu -- extern "C" entry points of the wide-feature path (include/pfs_b200.h, "Wide-feature path").
// Host-side only: argument checks, TMA tensor maps, grid sizing, launches on the caller's stream.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "../../include/pfs_b200.h"
#include "wide_gemm.cuh"
#include "wide_ops.cuh"

namespace pfs_host {
int fail_msg(int code, const char* msg);
void mark_launch(const char* name, cudaStream_t st);
int sm_count();
}  // namespace pfs_host

using namespace pfs;

namespace {

int wfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return pfs_host::fail_msg(code, buf);
}

#define W_REQUIRE(cond, msg)                                        \
    do {                                                            \
        if (!(cond)) return wfail(PFS_ERR_ARG, "%s (%s)", msg, #cond); \
    } while (0)
#define W_LAUNCH_CHECK(name)                                                                           \
    do {                                                                                              \
        cudaError_t e__ = cudaGetLastError();                                                         \
        if (e__ != cudaSuccess) return wfail(PFS_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e__)); \
        pfs_host::mark_launch(name, st);                                                              \
    } while (0)
#define W_TRY(expr)                      \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != PFS_OK) return rc__; \
    } while (0)

// ---- TMA tensor maps (driver entry point resolved at run time: no link dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// bf16 matrix [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 columns], 128-byte swizzle
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return wfail(PFS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    if (((uintptr_t)base & 15) != 0 || (ld % 8) != 0)
        return wfail(PFS_ERR_ARG, "TMA operand must be 16-byte aligned with a leading dimension multiple of 8 (ld=%lld)", ld);
    if (rows < 1 || cols < 1) return wfail(PFS_ERR_ARG, "empty TMA operand (%lld x %lld)", rows, cols);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return wfail(PFS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld %lld", (int)r, rows, cols, ld);
    return PFS_OK;
}

// a row of zeros standing in for an absent gather table (the kernel then has no per-table branches)
constexpr int kZeroRowElems = 4096;
const float* zero_row() {
    static float* p = nullptr;               // one per process; device memory is zero-filled once
    if (!p) {
        if (cudaMalloc(&p, kZeroRowElems * sizeof(float)) != cudaSuccess) return nullptr;
        if (cudaMemset(p, 0, kZeroRowElems * sizeof(float)) != cudaSuccess) return nullptr;
    }
    return p;
}

template <class Kern>
int allow_smem(Kern kern, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return wfail(PFS_ERR_CUDA, "cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    return PFS_OK;
}

// profile label: with PFS_PROFILE_SHAPES=1 the GEMM launches are reported per shape (bench.py kernel table)
const char* nt_name(const pfs_wide_gemm_args& a, bool tables, bool mask, bool bstat, int bn) {
    static const bool by_shape = getenv("PFS_PROFILE_SHAPES") && getenv("PFS_PROFILE_SHAPES")[0] == '1';
    if (!by_shape) return "k_wide_gemm_nt";
    static std::map<std::string, std::string> names;
    char buf[96];
    snprintf(buf, sizeof(buf), "k_wide_gemm_nt[M%s,N%d,K%d%s%s%s%s,bn%d]", a.M >= (1 << 20) ? "=E" : "<E", a.N, a.K + a.K2,
             tables ? ",tab" : "", mask ? ",mask" : "", a.A2 ? ",a2" : "", bstat ? ",bstat" : "", bn);
    return names.emplace(buf, buf).first->second.c_str();
}

template <int BN, int STAGES, bool TABLES, bool MASK, int BT = 0>
int launch_nt_impl(const pfs_wide_gemm_args& a, const GemmEpilogue& ep, cudaStream_t st) {
    CUtensorMap tmA, tmB, tmA2;
    W_TRY(make_map(&tmA, a.A, a.M, a.K, a.lda, kGemmBM));
    W_TRY(make_map(&tmB, a.B, a.N, (long long)a.K + a.K2, a.ldb, BN));
    if (a.A2) W_TRY(make_map(&tmA2, a.A2, a.a2_mod, a.K2, a.lda2, kGemmBM));
    else tmA2 = tmA;
    constexpr bool BSTAT = BT > 0;
    auto kern = k_wide_gemm_nt<BN, STAGES, TABLES, MASK, BT>;
    constexpr size_t smem = GemmNtSmem<BN, STAGES, BT>::bytes;
    W_TRY(allow_smem(kern, smem));
    const int nt = (a.N + BN - 1) / BN;
    const long long tiles = (long long)((a.M + kGemmBM - 1) / kGemmBM) * nt;
    int grid = (int)(tiles < pfs_host::sm_count() ? tiles : pfs_host::sm_count());
    if (BSTAT) grid -= grid % nt;            // every tile of a CTA must share its n block
    kern<<<grid, kGemmNtThreads, smem, st>>>(tmA, tmB, tmA2, ep, (__nv_bfloat16*)a.out_bf16, (int)a.ldc, a.M, a.N, a.K);
    W_LAUNCH_CHECK(nt_name(a, TABLES, MASK, BSTAT, BN));
    return PFS_OK;
}

// B-stationary variants (weight tile resident in shared memory) for large-M layers:
//   K_total <= 128: 256-column tiles (two k-blocks of B = 64 KB), full-rate MMAs
//   K_total <= 256: 128-column tiles (four k-blocks = 64 KB)
template <int BN, int BT>
int launch_nt_bstat(const pfs_wide_gemm_args& a, const GemmEpilogue& ep, cudaStream_t st) {
    const bool tables = ep.tab0 != nullptr, mask = ep.mask != nullptr;
    if (tables && mask) return launch_nt_impl<BN, 4, true, true, BT>(a, ep, st);
    if (tables) return launch_nt_impl<BN, 4, true, false, BT>(a, ep, st);
    if (mask) return launch_nt_impl<BN, 4, false, true, BT>(a, ep, st);
    return launch_nt_impl<BN, 4, false, false, BT>(a, ep, st);
}
// 0: not applicable; else the tile width to use
int bstat_tile(const pfs_wide_gemm_args& a) {
    const int kb = (a.K + kGemmBK - 1) / kGemmBK + (a.K2 + kGemmBK - 1) / kGemmBK;
    const long long mt = (a.M + kGemmBM - 1) / kGemmBM;
    if (a.N < 128 || mt < 8LL * pfs_host::sm_count()) return 0;
    if (kb <= 2 && a.N > 128 && pfs_host::sm_count() % ((a.N + 255) / 256) == 0) return 256;
    if (kb <= 4 && pfs_host::sm_count() % ((a.N + 127) / 128) == 0) return 128;
    return 0;
}

// FULL = gathered tables / derivative mask in the epilogue; the plain variant carries no code for them
template <int BN, int STAGES>
int launch_nt(const pfs_wide_gemm_args& a, const GemmEpilogue& ep, cudaStream_t st) {
    const bool tables = ep.tab0 != nullptr, mask = ep.mask != nullptr;
    if (tables && mask) return launch_nt_impl<BN, STAGES, true, true>(a, ep, st);
    if (tables) return launch_nt_impl<BN, STAGES, true, false>(a, ep, st);
    if (mask) return launch_nt_impl<BN, STAGES, false, true>(a, ep, st);
    return launch_nt_impl<BN, STAGES, false, false>(a, ep, st);
}

int tn_bn(int Kx) { return Kx > 128 ? 256 : Kx > 64 ? 128 : 64; }
// split of the contraction rows: enough CTAs to fill the device, each a multiple of the 64-row box
void tn_split(long long E, int J, int Kx, int& splits, int& rows_per_split) {
    const int BN = tn_bn(Kx);
    const long long tiles = (long long)((J + kGemmBM - 1) / kGemmBM) * ((Kx + BN - 1) / BN);
    long long want = (2LL * pfs_host::sm_count() + tiles - 1) / tiles;
    const long long max_splits = (E + 4 * kGemmBK - 1) / (4 * kGemmBK);     // at least 256 rows per split
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    long long rps = (E + want - 1) / want;
    rps = (rps + kGemmBK - 1) / kGemmBK * kGemmBK;
    splits = (int)((E + rps - 1) / rps);
    if (splits < 1) splits = 1;
    rows_per_split = (int)rps;
}

template <int BN, int STAGES>
int launch_tn(const void* D, long long ldd, const void* X, long long ldx, long long E, int J, int Kx, float* partial,
              int splits, int rps, cudaStream_t st) {
    CUtensorMap tmD, tmX;
    W_TRY(make_map(&tmD, D, E, J, ldd, kGemmBK));
    W_TRY(make_map(&tmX, X, E, Kx, ldx, kGemmBK));
    auto kern = k_wide_gemm_tn<BN, STAGES>;
    constexpr size_t smem = GemmTnSmem<BN, STAGES>::bytes;
    W_TRY(allow_smem(kern, smem));
    const int tiles = ((J + kGemmBM - 1) / kGemmBM) * ((Kx + BN - 1) / BN);
    kern<<<dim3(tiles, splits), kGemmThreads, smem, st>>>(tmD, tmX, partial, (int)E, J, Kx, rps);
    W_LAUNCH_CHECK("k_wide_gemm_tn");
    return PFS_OK;
}

int grid_for(long long items, int per_block = 256, int waves = 8) {
    long long b = (items + per_block - 1) / per_block;
    const long long cap = (long long)waves * pfs_host::sm_count();
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

SegDesc make_seg(const pfs_wide_segments& s) { return SegDesc{s.mode, s.nseg, s.S, s.T, s.ptr, s.list}; }
int check_seg(const pfs_wide_segments* s) {
    W_REQUIRE(s && s->nseg >= 1 && s->mode >= 0 && s->mode <= 2, "bad segments");
    if (s->mode == 2) W_REQUIRE(s->ptr, "listed segments need ptr");
    else W_REQUIRE(s->S >= 1 && s->T >= 1, "dense segments need S and T");
    return PFS_OK;
}
int segsum_chunks(const pfs_wide_segments& s, int C) {
    // few long segments (classes): split them so that the grid fills the device
    const long long blocks = s.nseg;
    const int want = (int)((4LL * pfs_host::sm_count() + blocks - 1) / blocks);
    long long len = s.mode == 0 ? s.T : s.mode == 1 ? s.S : 0;
    if (s.mode == 2) return want > 1 ? (want > 64 ? 64 : want) : 1;     // lengths unknown on the host: bounded split
    int n = want;
    if (n > len / 32) n = (int)(len / 32);
    (void)C;
    return n < 1 ? 1 : n;
}

int colstats_rowblocks(long long R, int C) {
    const int colblocks = (C + 63) / 64;
    long long nrb = (8LL * pfs_host::sm_count() + colblocks - 1) / colblocks;
    const long long maxrb = (R + 63) / 64;
    if (nrb > maxrb) nrb = maxrb;
    return (int)(nrb < 1 ? 1 : nrb);
}

}  // namespace

extern "C" {

size_t pfs_sizeof_wide_gemm_args(void) { return sizeof(pfs_wide_gemm_args); }
size_t pfs_sizeof_wide_segments(void) { return sizeof(pfs_wide_segments); }

int pfs_wide_gemm_nt(const pfs_wide_gemm_args* a) {
    W_REQUIRE(a && a->A && a->B && (a->out_bf16 || a->out_f32), "null pointer");
    W_REQUIRE(a->M >= 1 && a->N >= 8 && a->K >= 8 && a->N % 8 == 0 && a->K % 8 == 0, "bad GEMM sizes (N, K multiples of 8)");
    if (a->tab0 && !a->idx0) W_REQUIRE(a->div0 >= 1, "tab0 needs idx0 or div0");
    if (a->tab1 && !a->idx1) W_REQUIRE(a->mod1 >= 1, "tab1 needs idx1 or mod1");
    if (a->mask) W_REQUIRE(a->ldmask % 8 == 0 && ((uintptr_t)a->mask & 15) == 0, "mask alignment");
    if (a->out_f32) W_REQUIRE(a->ldf % 4 == 0 && ((uintptr_t)a->out_f32 & 15) == 0, "fp32 output alignment");
    cudaStream_t st = (cudaStream_t)a->stream;
    pfs_host::mark_launch(nullptr, st);
    GemmEpilogue ep{};
    ep.bias = a->bias; ep.bias_rowscale = a->bias_rowscale;
    ep.tab0 = a->tab0; ep.idx0 = a->idx0; ep.div0 = a->div0 > 0 ? a->div0 : 1;
    ep.tab1 = a->tab1; ep.idx1 = a->idx1; ep.mod1 = a->mod1 > 0 ? a->mod1 : 1;
    ep.rows0 = 1 << 30;
    if (ep.tab0 || ep.tab1) {
        // a single table: the other one reads a row of zeros (dense addressing, one row)
        W_REQUIRE(a->N <= kZeroRowElems, "gather tables wider than the zero row");
        if (!ep.tab0) { ep.tab0 = zero_row(); ep.idx0 = nullptr; ep.div0 = 1 << 30; ep.rows0 = 1; }
        if (!ep.tab1) { ep.tab1 = zero_row(); ep.idx1 = nullptr; ep.mod1 = 1; }
        if (!ep.tab0 || !ep.tab1) return wfail(PFS_ERR_CUDA, "could not allocate the zero row");
        W_REQUIRE(a->N % 4 == 0 && ((uintptr_t)ep.tab0 & 15) == 0 && ((uintptr_t)ep.tab1 & 15) == 0, "gather table alignment");
    }
    ep.mask = (const __nv_bfloat16*)a->mask; ep.ldmask = (int)a->ldmask;
    ep.act = a->act;
    ep.brow = a->bias_rows; ep.brow_div = a->bias_rows_div > 0 ? a->bias_rows_div : 1;
    ep.k1 = 0; ep.a2_mod = 1;
    if (a->bias_rows) W_REQUIRE(a->bias_rows_div % kGemmBM == 0 && ((uintptr_t)a->bias_rows & 15) == 0,
                                "bias_rows: tiles of 128 rows must not straddle a row block (div % 128 == 0)");
    if (a->A2) {
        W_REQUIRE(a->K2 >= 8 && a->K2 % 8 == 0 && a->a2_mod >= kGemmBM && a->a2_mod % kGemmBM == 0,
                  "A2: K2 multiple of 8, a2_mod a multiple of 128 (dense layout with T % 128 == 0)");
        ep.k1 = a->K2; ep.a2_mod = a->a2_mod;
    }
    ep.out_f32 = a->out_f32; ep.ldf = (int)a->ldf;
    ep.out_bf16 = a->out_bf16 ? 1 : 0;
    if (a->out_bf16) W_REQUIRE(a->ldc % 4 == 0 && ((uintptr_t)a->out_bf16 & 7) == 0, "bf16 output alignment");
    const int bs = bstat_tile(*a);
    if (bs == 256) return launch_nt_bstat<256, 2>(*a, ep, st);
    if (bs == 128) return launch_nt_bstat<128, 4>(*a, ep, st);
    if (a->N > 128) return launch_nt<256, 3>(*a, ep, st);
    if (a->N > 64) return launch_nt<128, 5>(*a, ep, st);
    return launch_nt<64, 6>(*a, ep, st);
}

size_t pfs_wide_gemm_tn_workspace(int64_t E, int32_t J, int32_t Kx) {
    int splits, rps;
    tn_split(E, J, Kx, splits, rps);
    return (size_t)splits * J * Kx * sizeof(float) + 256;
}

int pfs_wide_gemm_tn(const void* D, int64_t ldd, const void* X, int64_t ldx, int64_t E, int32_t J, int32_t Kx,
                     float* out, int64_t ldo, int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    W_REQUIRE(D && X && out && workspace, "null pointer");
    W_REQUIRE(E >= 1 && E < (1ll << 31) && J >= 8 && Kx >= 8 && J % 8 == 0 && Kx % 8 == 0, "bad sizes (J, Kx multiples of 8)");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    int splits, rps;
    tn_split(E, J, Kx, splits, rps);
    if (workspace_bytes < (size_t)splits * J * Kx * sizeof(float)) return wfail(PFS_ERR_WORKSPACE, "gemm_tn: workspace too small");
    float* partial = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const int BN = tn_bn(Kx);
    if (BN == 256) W_TRY((launch_tn<256, 4>(D, ldd, X, ldx, E, J, Kx, partial, splits, rps, st)));
    else if (BN == 128) W_TRY((launch_tn<128, 6>(D, ldd, X, ldx, E, J, Kx, partial, splits, rps, st)));
    else W_TRY((launch_tn<64, 8>(D, ldd, X, ldx, E, J, Kx, partial, splits, rps, st)));
    const int n = J * Kx;
    k_wide_reduce_splits<<<(n + 255) / 256, 256, 0, st>>>(partial, splits, J, Kx, out, (int)ldo, 0, accumulate);
    W_LAUNCH_CHECK("k_wide_reduce_splits");
    return PFS_OK;
}

size_t pfs_wide_colstats_workspace(int64_t R, int32_t C) {
    return (size_t)colstats_rowblocks(R, C) * 2 * C * sizeof(float) + 256;
}

int pfs_wide_colstats(int32_t kind, const void* g, int32_t g_dtype, int64_t ldg, const void* v, int32_t v_dtype,
                      int64_t ldv, const float* p0, const float* p1, const float* roww, int64_t R, int32_t C,
                      float* out, void* workspace, size_t workspace_bytes, void* stream) {
    W_REQUIRE(g && out && workspace && R >= 1 && C >= 2 && C % 2 == 0, "bad arguments");
    W_REQUIRE(kind == 0 || kind == 1, "kind");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    const int nrb = colstats_rowblocks(R, C);
    if (workspace_bytes < (size_t)nrb * 2 * C * sizeof(float)) return wfail(PFS_ERR_WORKSPACE, "colstats: workspace too small");
    float* partial = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const long long rpb = (R + nrb - 1) / nrb;
    const dim3 grid((C + 63) / 64, nrb), block(32, 8);
#define PFS_COLSTATS(TG, TV)                                                                                         \
    k_wide_colstats<TG, TV><<<grid, block, 0, st>>>(kind, (const TG*)g, (int)ldg, (const TV*)v, (int)ldv, p0, p1, roww, R, C, \
                                                    rpb, partial)
    if (g_dtype == 0 && (v_dtype == 0 || !v)) PFS_COLSTATS(bf16, bf16);
    else if (g_dtype == 0 && v_dtype == 1) PFS_COLSTATS(bf16, float);
    else if (g_dtype == 1 && (v_dtype == 1 || !v)) PFS_COLSTATS(float, float);
    else if (g_dtype == 1 && v_dtype == 0) PFS_COLSTATS(float, bf16);
    else return wfail(PFS_ERR_ARG, "colstats: bad dtype codes");
#undef PFS_COLSTATS
    W_LAUNCH_CHECK("k_wide_colstats");
    if (g_dtype == 0) k_wide_colstats_final<bf16><<<(C + 127) / 128, 128, 0, st>>>(kind, partial, nrb, C, R, (const bf16*)g, out);
    else k_wide_colstats_final<float><<<(C + 127) / 128, 128, 0, st>>>(kind, partial, nrb, C, R, (const float*)g, out);
    W_LAUNCH_CHECK("k_wide_colstats_final");
    return PFS_OK;
}

int pfs_wide_rowmap(int32_t kind, const void* x, int32_t x_dtype, int64_t ldx, const void* v, int32_t v_dtype, int64_t ldv,
                    const float* a, const float* b, const float* p0, const float* p1, const float* c2, int64_t R, int32_t C,
                    void* out_bf16, int64_t ldo, void* stream) {
    W_REQUIRE(x && a && b && out_bf16 && R >= 1 && C >= 2 && C % 2 == 0, "bad arguments");
    if (kind == 1) W_REQUIRE(v && p0 && p1 && c2, "kind 1 needs v, p0, p1, c2");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    // 8 columns per thread (16-byte accesses, coefficients in registers) when the rows allow it
    const bool vec = C % 8 == 0 && C <= 2048 && ldx % 8 == 0 && ldo % 8 == 0 && (!v || ldv % 8 == 0) &&
                     ((uintptr_t)x & 15) == 0 && ((uintptr_t)out_bf16 & 15) == 0 && (!v || ((uintptr_t)v & 15) == 0) &&
                     ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0;
    const int grid = vec ? grid_for(R, 256 / (C / 8 > 256 ? 256 : C / 8)) : grid_for(R * (C / 2));
#define PFS_ROWMAP(TX, TV)                                                                                            \
    if (vec)                                                                                                          \
        k_wide_rowmap8<TX, TV><<<grid, 256, 0, st>>>(kind, (const TX*)x, (int)ldx, (const TV*)v, (int)ldv, a, b, p0, p1, c2, R, \
                                                     C, (bf16*)out_bf16, (int)ldo);                                   \
    else                                                                                                              \
        k_wide_rowmap<TX, TV><<<grid, 256, 0, st>>>(kind, (const TX*)x, (int)ldx, (const TV*)v, (int)ldv, a, b, p0, p1, c2, R, \
                                                    C, (bf16*)out_bf16, (int)ldo)
    if (x_dtype == 0 && (v_dtype == 0 || !v)) PFS_ROWMAP(bf16, bf16);
    else if (x_dtype == 0 && v_dtype == 1) PFS_ROWMAP(bf16, float);
    else if (x_dtype == 1 && (v_dtype == 1 || !v)) PFS_ROWMAP(float, float);
    else if (x_dtype == 1 && v_dtype == 0) PFS_ROWMAP(float, bf16);
    else return wfail(PFS_ERR_ARG, "rowmap: bad dtype codes");
#undef PFS_ROWMAP
    W_LAUNCH_CHECK("k_wide_rowmap");
    return PFS_OK;
}

size_t pfs_wide_segsum_workspace(const pfs_wide_segments* sd, int32_t C) {
    if (!sd) return 0;
    return (size_t)segsum_chunks(*sd, C) * sd->nseg * C * sizeof(float) + 256;
}

int pfs_wide_segsum(const pfs_wide_segments* sd, const void* x, int32_t x_dtype, int64_t ldx, int32_t C, float* out_f32,
                    void* out_bf16, void* workspace, size_t workspace_bytes, void* stream) {
    W_TRY(check_seg(sd));
    W_REQUIRE(x && (out_f32 || out_bf16) && C >= 8 && C % 8 == 0 && ldx % 8 == 0 && ((uintptr_t)x & 15) == 0,
              "bad arguments (C and ldx multiples of 8, 16-byte aligned rows)");
    W_REQUIRE(x_dtype == 0 || x_dtype == 1, "dtype code (0 = bf16, 1 = fp32)");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    const int nchunk = segsum_chunks(*sd, C);
    float* partial = nullptr;
    if (nchunk > 1) {
        if (!workspace || workspace_bytes < (size_t)nchunk * sd->nseg * C * sizeof(float))
            return wfail(PFS_ERR_WORKSPACE, "segsum: workspace too small");
        partial = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    }
    if (x_dtype == 0)
        k_wide_segsum<bf16><<<dim3(sd->nseg, nchunk), 256, 0, st>>>(make_seg(*sd), (const bf16*)x, (int)ldx, C, nchunk, out_f32,
                                                                    (bf16*)out_bf16, partial);
    else
        k_wide_segsum<float><<<dim3(sd->nseg, nchunk), 256, 0, st>>>(make_seg(*sd), (const float*)x, (int)ldx, C, nchunk, out_f32,
                                                                     (bf16*)out_bf16, partial);
    W_LAUNCH_CHECK("k_wide_segsum");
    if (nchunk > 1) {
        const long long n = (long long)sd->nseg * C;
        k_wide_segsum_final<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, nchunk, n, out_f32, (bf16*)out_bf16);
        W_LAUNCH_CHECK("k_wide_segsum_final");
    }
    return PFS_OK;
}

int pfs_wide_moments_fwd(const pfs_wide_segments* sd, const void* m, int32_t m_dtype, int32_t C, float* moments, void* stream) {
    W_TRY(check_seg(sd));
    W_REQUIRE(m && moments && C >= 8 && C % 8 == 0 && C <= 2048 && ((uintptr_t)m & 15) == 0, "bad arguments");
    W_REQUIRE(m_dtype == 0 || m_dtype == 1, "dtype code (0 = bf16, 1 = fp32)");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    if (m_dtype == 0) k_wide_moments_fwd<bf16><<<sd->nseg, 256, 0, st>>>(make_seg(*sd), (const bf16*)m, C, moments);
    else k_wide_moments_fwd<float><<<sd->nseg, 256, 0, st>>>(make_seg(*sd), (const float*)m, C, moments);
    W_LAUNCH_CHECK("k_wide_moments_fwd");
    return PFS_OK;
}

int pfs_wide_source_hcat(const void* x_s_bf16, const float* moments, int32_t S, int32_t F, void* hcat_bf16, int64_t ldo,
                         int32_t with_lo, void* stream) {
    W_REQUIRE(x_s_bf16 && moments && hcat_bf16 && S >= 1 && F >= 2, "bad arguments");
    W_REQUIRE(ldo >= (with_lo ? 17 : 9) * (int64_t)F, "hcat rows hold 9F columns (17F with the remainders)");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    k_wide_source_hcat<<<grid_for((long long)S * 3 * F), 256, 0, st>>>((const bf16*)x_s_bf16, moments, S, F, (bf16*)hcat_bf16,
                                                                      (int)ldo, with_lo);
    W_LAUNCH_CHECK("k_wide_source_hcat");
    return PFS_OK;
}

int pfs_wide_split(const float* x, int64_t ldx, int64_t R, int32_t C, void* out_bf16, int64_t ldo, void* stream) {
    W_REQUIRE(x && out_bf16 && R >= 1 && C >= 2 && C % 2 == 0 && ldx % 2 == 0 && ldo % 2 == 0 && ldo >= 2 * (int64_t)C &&
                  ((uintptr_t)x & 7) == 0 && ((uintptr_t)out_bf16 & 3) == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    k_wide_split<<<grid_for(R * (C / 2)), 256, 0, st>>>(x, (int)ldx, R, C, (bf16*)out_bf16, (int)ldo);
    W_LAUNCH_CHECK("k_wide_split");
    return PFS_OK;
}

int pfs_wide_source_coef(const pfs_wide_segments* sd, const float* dh, const float* moments, int32_t S, int32_t F,
                         void* dx_s_bf16, float* coef, void* stream) {
    W_TRY(check_seg(sd));
    W_REQUIRE(dh && moments && dx_s_bf16 && coef && S >= 1 && F >= 2 && sd->nseg == S, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    k_wide_source_coef<<<grid_for((long long)S * 3 * F), 256, 0, st>>>(make_seg(*sd), dh, moments, S, F, (bf16*)dx_s_bf16, coef);
    W_LAUNCH_CHECK("k_wide_source_coef");
    return PFS_OK;
}

int pfs_wide_source_dm(const void* m, int32_t m_dtype, const float* moments, const float* coef, const int32_t* src, int32_t T,
                       int64_t E, int32_t C, void* dm_bf16, void* stream) {
    W_REQUIRE(m && moments && coef && dm_bf16 && E >= 1 && C >= 8 && C % 8 == 0 && (src || T >= 1), "bad arguments");
    W_REQUIRE(m_dtype == 0 || m_dtype == 1, "dtype code (0 = bf16, 1 = fp32)");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    if (m_dtype == 0)
        k_wide_source_dm<bf16><<<grid_for(E * (C / 8)), 256, 0, st>>>((const bf16*)m, moments, coef, src, T, E, C, (bf16*)dm_bf16);
    else
        k_wide_source_dm<float><<<grid_for(E * (C / 8)), 256, 0, st>>>((const float*)m, moments, coef, src, T, E, C, (bf16*)dm_bf16);
    W_LAUNCH_CHECK("k_wide_source_dm");
    return PFS_OK;
}

int pfs_wide_source_dm_seg(const pfs_wide_segments* sd, const void* m, int32_t m_dtype, const float* moments, const float* coef,
                           int32_t C, void* dm_bf16, void* stream) {
    W_TRY(check_seg(sd));
    W_REQUIRE(m && moments && coef && dm_bf16 && C >= 8 && C % 8 == 0 && C <= 2048, "bad arguments");
    W_REQUIRE(m_dtype == 0 || m_dtype == 1, "dtype code (0 = bf16, 1 = fp32)");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    if (m_dtype == 0) k_wide_source_dm_seg<bf16><<<sd->nseg, 256, 0, st>>>(make_seg(*sd), (const bf16*)m, moments, coef, C, (bf16*)dm_bf16);
    else k_wide_source_dm_seg<float><<<sd->nseg, 256, 0, st>>>(make_seg(*sd), (const float*)m, moments, coef, C, (bf16*)dm_bf16);
    W_LAUNCH_CHECK("k_wide_source_dm");
    return PFS_OK;
}

int pfs_wide_gather_mask(const float* tab, const int32_t* idx, int32_t mod, const void* act_bf16, int64_t E, int32_t C,
                         void* out_bf16, void* stream) {
    W_REQUIRE(tab && act_bf16 && out_bf16 && E >= 1 && C >= 8 && C % 8 == 0 && (idx || mod >= 1), "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    k_wide_gather_mask<<<grid_for(E * (C / 8)), 256, 0, st>>>(tab, idx, mod, (const bf16*)act_bf16, E, C, (bf16*)out_bf16);
    W_LAUNCH_CHECK("k_wide_gather_mask");
    return PFS_OK;
}

int pfs_wide_head_fwd(const void* a_bf16, const float* w2, const float* b2, float scale, int64_t E, int32_t F,
                      const float* class_hours, const int32_t* tgt, int32_t T, float* pred, float* time, float* visits,
                      float* time_int, void* stream) {
    W_REQUIRE(a_bf16 && w2 && b2 && pred && time && E >= 1, "bad arguments");
    W_REQUIRE(F >= 8 && F % 8 == 0 && F <= 256 && ((F >> 3) & ((F >> 3) - 1)) == 0, "time head: Fdim must be 8 * a power of two, <= 256");
    if (visits || time_int) W_REQUIRE(visits && time_int && class_hours && (tgt || T >= 1), "integer outputs need hours and classes");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    const int rows_per_block = 8 * (32 / (F >> 3));
    k_wide_head_fwd<<<grid_for(E, rows_per_block), 256, 0, st>>>((const bf16*)a_bf16, w2, b2, scale, E, F, class_hours, tgt, T,
                                                                  pred, time, visits, time_int);
    W_LAUNCH_CHECK("k_wide_head_fwd");
    return PFS_OK;
}

int pfs_wide_head_bwd(const void* a_bf16, const float* w2, const float* pred, const float* g_time, float scale, int64_t E,
                      int32_t F, float* gp, void* da_bf16, void* stream) {
    W_REQUIRE(a_bf16 && w2 && pred && g_time && gp && da_bf16 && E >= 1 && F >= 8 && F % 8 == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    k_wide_head_bwd<<<grid_for(E * (F / 8)), 256, 0, st>>>((const bf16*)a_bf16, w2, pred, g_time, scale, E, F, gp, (bf16*)da_bf16);
    W_LAUNCH_CHECK("k_wide_head_bwd");
    return PFS_OK;
}

int pfs_wide_cast(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t n, void* stream) {
    W_REQUIRE(in && out && n >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    const int grid = grid_for(n);
    if (in_dtype == 0 && out_dtype == 1) k_wide_cast<bf16, float><<<grid, 256, 0, st>>>((const bf16*)in, n, (float*)out);
    else if (in_dtype == 1 && out_dtype == 0) k_wide_cast<float, bf16><<<grid, 256, 0, st>>>((const float*)in, n, (bf16*)out);
    else return wfail(PFS_ERR_ARG, "cast: dtype codes must differ (0 = bf16, 1 = fp32)");
    W_LAUNCH_CHECK("k_wide_cast");
    return PFS_OK;
}

int pfs_wide_transpose(const void* in_bf16, int32_t R, int32_t C, int64_t ld, void* out_bf16, void* stream) {
    W_REQUIRE(in_bf16 && out_bf16 && R >= 1 && C >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    pfs_host::mark_launch(nullptr, st);
    k_wide_transpose<<<dim3((C + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, st>>>((const bf16*)in_bf16, R, C, (int)ld, (bf16*)out_bf16);
    W_LAUNCH_CHECK("k_wide_transpose");
    return PFS_OK;
}

}  // extern "C"
