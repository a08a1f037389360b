// probe: which shared-memory word does tcgen05.mma kind::tf32 fetch for an MN-major no-swizzle operand?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pfs-neural-net_b200/csrc/tc_ptx.cuh"
using namespace pfs;
constexpr int WORDS = 12288;   // 48 KB operand under test
// test 0: A MN-major (probed), B K-major identity N=16 ; test 1: A K-major identity, B MN-major (probed) N=NB
__global__ void probe(int test, int fill, uint32_t lbo, uint32_t sbo, int NB, float* out) {
    extern __shared__ __align__(1024) float sm[];
    float* X = sm;                  // probed operand
    float* I = sm + WORDS;          // identity operand, K-major: (k/4)*2048 + row*16 + (k%4)*4, 128 rows
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) X[i] = fill == 0 ? (float)(i % 1024) : (float)(i / 1024);
    for (int i = threadIdx.x; i < 2 * 512; i += blockDim.x) I[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < 8) { int k = threadIdx.x; I[((k / 4) * 2048 + k * 16 + (k % 4) * 4) / 4] = 1.f; }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    if (threadIdx.x < 32) tmem_alloc(&slot, 128);
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tm = slot;
    int N = (test == 0 || test == 2) ? 16 : NB;
    if (threadIdx.x == 0) {
        uint64_t dx = umma_desc_ls(smem_u32(X), lbo, sbo), di = umma_desc_ls(smem_u32(I), 2048, 128);
        if (test == 0) umma_tf32(tm, dx, di, umma_idesc_tf32_mn(128, N, 1, 0), 0);
        else if (test == 1) umma_tf32(tm, di, dx, umma_idesc_tf32_mn(128, N, 0, 1), 0);
        else if (test == 2) umma_tf32(tm, dx, di, umma_idesc_tf32_mn(128, N, 0, 0), 0);
        else umma_tf32(tm, di, dx, umma_idesc_tf32_mn(128, N, 0, 0), 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0); tc_fence_after();
    int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp < 4) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int q = 0; q < 16; ++q) out[(warp * 32 + lane) * 128 + c0 + q] = v[q];
        }
    }
    if (threadIdx.x == 0) { out[127 * 128 + 127] = 777.f; out[127 * 128 + 126] = X[5]; out[127 * 128 + 125] = I[0]; out[127*128+124] = (float)tm; }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 128);
}
int main() {
    float* d; cudaMalloc(&d, 128 * 128 * 4);
    cudaError_t ea = cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (WORDS + 1024 + 2048) * 4);
    printf("attr: %s\n", cudaGetErrorString(ea));
    std::vector<float> h0(128 * 128), h1(128 * 128);
    uint32_t cfg[][2] = {{128, 1040}, {1040, 128}, {128, 2048}, {2048, 128}};
    for (auto& c : cfg) for (int test = 0; test < 4; ++test) {
        int NB = 96;
        for (int fill = 0; fill < 2; ++fill) {
            cudaMemset(d, 0, 128 * 128 * 4);
            probe<<<1, 128, (WORDS + 1024 + 2048) * 4>>>(test, fill, c[0], c[1], NB, d);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return 1; }
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(fill == 0 ? h0.data() : h1.data(), d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
        }
        printf("sentinels %g %g %g %g\n", h0[127*128+127], h0[127*128+126], h0[127*128+125], h0[127*128+124]);
        printf("== lbo=%u sbo=%u test=%d (%s MN-major probed)\n", c[0], c[1], test, test == 0 ? "A" : "B");
        // test 0: D[m][n<8] = A(m, k=n); test 1: D[m<8][n] = B(n, k=m)
        int MN = (test == 0 || test == 2) ? 128 : NB;
        for (int mn = 0; mn < MN; mn += (mn < 10 ? 1 : (mn < 40 ? 6 : 24))) {
            printf("  mn=%3d byte offsets for k=0..7:", mn);
            for (int k = 0; k < 8; ++k) {
                int idx = (test == 0 || test == 2) ? mn * 128 + k : k * 128 + mn;
                int w = (int)h1[idx] * 1024 + (int)h0[idx];
                printf(" %6d", w * 4);
            }
            printf("\n");
        }
    }
    return 0;
}
