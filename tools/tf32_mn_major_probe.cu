// probe 2: does an MN-major no-swizzle tf32 MMA execute at all?  Prefill D with a K-major MMA, then run variants.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pfs-neural-net_b200/csrc/tc_ptx.cuh"
using namespace pfs;
constexpr int WORDS = 12288;
__global__ void probe(int a_mn, int b_mn, int both_x, uint32_t lbo, uint32_t sbo, int prefill, float* out) {
    extern __shared__ __align__(1024) float sm[];
    float* X = sm;
    float* I = sm + WORDS;          // identity, valid for K-major (lbo 2048, sbo 128) and, for rows < 8, MN-major (sbo 2048)
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) X[i] = (float)(i % 1024);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) I[i] = 0.f;
    __syncthreads();
    // symmetric identity: K-major element (r,k) at (k/4)*2048 + r*16 + (k%4)*4 ; MN-major (assumed) element (n,k) at (n/4)*2048 + k*16 + (n%4)*4
    if (threadIdx.x < 8) { int k = threadIdx.x; I[((k / 4) * 2048 + k * 16 + (k % 4) * 4) / 4] = 1.f; }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    if (threadIdx.x < 32) tmem_alloc(&slot, 128);
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tm = slot;
    const int N = 16;
    if (threadIdx.x == 0) {
        uint64_t dxk = umma_desc_ls(smem_u32(X), 128, 1040), dik = umma_desc_ls(smem_u32(I), 2048, 128);
        uint64_t dxm = umma_desc_ls(smem_u32(X), lbo, sbo), dim_ = umma_desc_ls(smem_u32(I), 128, 2048);
        if (prefill) umma_tf32(tm, dxk, dik, umma_idesc_tf32_mn(128, N, 0, 0), 0);     // D[m][n<8] = word index of K-major fetch
        uint64_t da = a_mn ? dxm : dxk;
        uint64_t db = both_x ? (b_mn ? dxm : dxk) : (b_mn ? dim_ : dik);
        umma_tf32(tm, da, db, umma_idesc_tf32_mn(128, N, a_mn, b_mn), prefill == 2 ? 1 : 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0); tc_fence_after();
    int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp < 4) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), v);
        for (int q = 0; q < 16; ++q) out[(warp * 32 + lane) * 16 + q] = v[q];
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 128);
}
int main() {
    float* d; cudaMalloc(&d, 128 * 16 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (WORDS + 4096) * 4);
    std::vector<float> h(128 * 16);
    struct Cfg { int a_mn, b_mn, both_x; uint32_t lbo, sbo; int prefill; };
    Cfg cfgs[] = {{0,0,0,128,1040,0}, {1,0,0,128,1040,0}, {1,0,0,128,1040,1}, {1,0,0,128,1040,2}, {0,1,0,128,1040,1}, {1,1,0,128,1040,0}, {1,1,0,128,1040,1},
                  {1,1,0,1040,128,0}, {1,0,0,1040,128,0}, {1,1,1,128,1040,0}};
    for (auto& c : cfgs) {
        cudaMemset(d, 0, 128 * 16 * 4);
        probe<<<1, 128, (WORDS + 4096) * 4>>>(c.a_mn, c.b_mn, c.both_x, c.lbo, c.sbo, c.prefill, d);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 128 * 16 * 4, cudaMemcpyDeviceToHost);
        printf("== a_mn=%d b_mn=%d both_x=%d lbo=%u sbo=%u prefill=%d\n", c.a_mn, c.b_mn, c.both_x, c.lbo, c.sbo, c.prefill);
        for (int m : {0, 1, 2, 3, 4, 5, 8, 9, 17, 40}) {
            printf("  m=%3d:", m);
            for (int n = 0; n < 10; ++n) printf(" %8g", h[m * 16 + n]);
            printf("\n");
        }
    }
    return 0;
}
