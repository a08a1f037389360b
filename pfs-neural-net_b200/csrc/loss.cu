// loss.cu -- the training loss that follows the message-passing path (reference src/train.py:21-80;
// SURVEY.md section 8f row N1): softfloor, class / fibre sums, completeness, penalties, variance and the
// gradient w.r.t. the edge times, as five small kernels instead of ~45 ATen ops and two torch_scatter
// calls.  fp32, dense canonical edge order (the reference reshapes the times to [NFIBERS, NCLASSES],
// src/train.py:67), deterministic: every reduction has a fixed order.
#include <cstdarg>
#include <cstdio>
#include <math_constants.h>

#include <cuda_runtime.h>

#include "../../include/pfs_b200.h"

namespace pfs_host {
int fail_msg(int code, const char* msg);
void mark_launch(const char* name, cudaStream_t st);
int sm_count();
}  // namespace pfs_host

namespace {

int lfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return pfs_host::fail_msg(code, buf);
}
#define L_REQUIRE(cond, msg)                                         \
    do {                                                             \
        if (!(cond)) return lfail(PFS_ERR_ARG, "%s (%s)", msg, #cond); \
    } while (0)
#define L_LAUNCH_CHECK(name)                                                                           \
    do {                                                                                              \
        cudaError_t e__ = cudaGetLastError();                                                         \
        if (e__ != cudaSuccess) return lfail(PFS_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e__)); \
        pfs_host::mark_launch(name, st);                                                              \
    } while (0)

struct LossConst {
    int S, T;
    float total_time, wutils, wvar, pclass, pfiber, noiselevel;
    float r, atan_r;      // r = exp(-1 / sharpness) (0 when sharpness == 0), atan(r / (1 - r))
    const float* sharp_dev;   // optional device scalar overriding the sharpness (CUDA-graph replays with a schedule)
};
__device__ __forceinline__ void resolve_sharpness(LossConst& c) {
    if (c.sharp_dev) {
        const float sh = *c.sharp_dev;
        c.r = sh == 0.f ? 0.f : expf(-1.f / sh);
        c.atan_r = atanf(c.r / (1.f - c.r));
    }
}

// softfloor (reference src/train.py:21-27) of the noisy visit count and its derivative
__device__ __forceinline__ void softfloor_eval(float visited, float u, const LossConst& c, float& sf, float& dsf) {
    const float x = visited + c.noiselevel * (u - 0.5f);
    float sn, cs;
    sincosf(2.f * CUDART_PI_F * x, &sn, &cs);
    const float den = 1.f - c.r * cs;
    sf = x + (atanf(c.r * sn / den) - c.atan_r) * (1.f / CUDART_PI_F);
    dsf = 1.f + 2.f * c.r * (cs - c.r) / (1.f - 2.f * c.r * cs + c.r * c.r);
}

// per edge: galaxies = max(0, softfloor(time / T_i)), time2 = galaxies * T_i      (src/train.py:43-49)
__global__ void __launch_bounds__(256) k_loss_edge_fwd(LossConst c, const float* __restrict__ time,
                                                       const float* __restrict__ noise, const float* __restrict__ hours,
                                                       float* __restrict__ galaxies, float* __restrict__ time2) {
    resolve_sharpness(c);
    const long long E = (long long)c.S * c.T;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const float h = hours[e % c.T];
        float sf, dsf;
        softfloor_eval(time[e] / h, noise[e], c, sf, dsf);
        const float g = fmaxf(sf, 0.f);
        galaxies[e] = g;
        time2[e] = g * h;
    }
}

// fibre sums of time2 (src/train.py:61): one warp per fibre
__global__ void __launch_bounds__(256) k_loss_fibre_sums(int S, int T, const float* __restrict__ time2, float* __restrict__ fibre_time) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= S) return;
    float s = 0.f;
    for (int i = lane; i < T; i += 32) s += time2[(size_t)k * T + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) fibre_time[k] = s;
}

// class statistics over the fibres: partial[rb][0][i] = sum galaxies, [1] = sum (time2 - shift_i), [2] = sum (time2 - shift_i)^2,
// shift_i = time2 of fibre 0 (shifted sums do not cancel).  grid = row blocks, 256 threads = (row lane, class)
__global__ void __launch_bounds__(256) k_loss_class_partial(int S, int T, const float* __restrict__ galaxies,
                                                            const float* __restrict__ time2, int rows_per_block,
                                                            float* __restrict__ partial) {
    __shared__ float red[3][256];
    const int tc = T < 256 ? T : 256;
    const int lanes = 256 / tc;
    const int lr = threadIdx.x / tc, lc = threadIdx.x - lr * tc;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(S, r0 + rows_per_block);
    for (int c0 = 0; c0 < T; c0 += tc) {
        const int i = c0 + lc;
        float a = 0.f, b = 0.f, q = 0.f;
        if (lr < lanes && i < T) {
            const float shift = time2[i];
            for (int k = r0 + lr; k < r1; k += lanes) {
                const size_t e = (size_t)k * T + i;
                a += galaxies[e];
                const float d = time2[e] - shift;
                b += d;
                q = fmaf(d, d, q);
            }
        }
        __syncthreads();
        red[0][threadIdx.x] = a; red[1][threadIdx.x] = b; red[2][threadIdx.x] = q;
        __syncthreads();
        if (lr == 0 && i < T) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            for (int r = 0; r < lanes; ++r) {
                s0 += red[0][r * tc + lc]; s1 += red[1][r * tc + lc]; s2 += red[2][r * tc + lc];
            }
            float* o = partial + (size_t)blockIdx.x * 3 * T;
            o[i] = s0; o[T + i] = s1; o[2 * T + i] = s2;
        }
    }
}

__device__ __forceinline__ double block_sum(double v, double* red) {     // fixed tree, blockDim.x == 1024
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

// one CTA: class totals, completeness and its minimum, penalties, variance, loss (src/train.py:51-71) and the per-class
// gradient coefficient dL/dn'_i.  scalars = {loss, totutils, class_penalty, fibre_penalty, variance, #minima}
__global__ void __launch_bounds__(1024) k_loss_scalars(const LossConst c, const float* __restrict__ partial, int nrb,
                                                       const float* __restrict__ counts, const float* __restrict__ time2,
                                                       const float* __restrict__ fibre_time, float* __restrict__ n_prime,
                                                       float* __restrict__ class_mean, float* __restrict__ class_coef,
                                                       float* __restrict__ scalars) {
    __shared__ double red[1024];
    __shared__ float s_min;
    const int S = c.S, T = c.T;
    double var_sum = 0.0, cpen = 0.0;
    float my_min = CUDART_INF_F;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int b = 0; b < nrb; ++b) {
            s0 += partial[(size_t)b * 3 * T + i];
            s1 += partial[(size_t)b * 3 * T + T + i];
            s2 += partial[(size_t)b * 3 * T + 2 * T + i];
        }
        const double n = (double)S;
        n_prime[i] = (float)s0;
        class_mean[i] = (float)((double)time2[i] + s1 / n);
        const double m2 = s2 - s1 * s1 / n;
        var_sum += (m2 > 0.0 ? m2 : 0.0) / (n - 1.0);                 // torch.var: unbiased
        const double over = s0 - (double)counts[i];
        if (over > 0.0) cpen += over * over;
        my_min = fminf(my_min, (float)s0 / counts[i]);
    }
    // minimum completeness and the number of classes attaining it (torch.min spreads the gradient evenly over ties)
    red[threadIdx.x] = (double)my_min;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) s_min = (float)red[0];
    __syncthreads();
    const float cmin = s_min;
    double ties = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x)
        if (n_prime[i] / counts[i] == cmin) ties += 1.0;
    ties = block_sum(ties, red);
    var_sum = block_sum(var_sum, red);
    cpen = block_sum(cpen, red);
    double fpen = 0.0;
    for (int k = threadIdx.x; k < S; k += blockDim.x) {
        const double ot = (double)fibre_time[k] - (double)c.total_time;
        const double l = ot > 0.0 ? ot : 0.1 * ot;                     // nn.LeakyReLU(0.1), src/train.py:63
        fpen += l * l;
    }
    fpen = block_sum(fpen, red);
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const float np = n_prime[i], N = counts[i];
        float coef = 2.f * c.pclass * fmaxf(np - N, 0.f);
        if (np / N == cmin) coef -= c.wutils / ((float)ties * N);
        class_coef[i] = coef;
    }
    if (threadIdx.x == 0) {
        const double class_penalty = c.pclass * cpen, fibre_penalty = c.pfiber * fpen;
        scalars[0] = (float)(-(double)c.wutils * cmin + fibre_penalty + class_penalty - (double)c.wvar * var_sum);
        scalars[1] = cmin;
        scalars[2] = (float)class_penalty;
        scalars[3] = (float)fibre_penalty;
        scalars[4] = (float)var_sum;
        scalars[5] = (float)ties;
    }
}

// g_time[e] = gL * [softfloor > 0] * softfloor' / T_i * (dL/dn'_i + T_i * (dL/dfibre_time_k - 2 wvar (time2 - mean_i) / (S - 1)))
__global__ void __launch_bounds__(256) k_loss_edge_bwd(LossConst c, const float* __restrict__ time,
                                                       const float* __restrict__ noise, const float* __restrict__ hours,
                                                       const float* __restrict__ time2, const float* __restrict__ fibre_time,
                                                       const float* __restrict__ class_mean, const float* __restrict__ class_coef,
                                                       const float* __restrict__ g_loss, float* __restrict__ g_time) {
    resolve_sharpness(c);
    const long long E = (long long)c.S * c.T;
    const float gl = g_loss ? g_loss[0] : 1.f;
    const float vscale = 2.f * c.wvar / (float)(c.S - 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % c.T);
        const long long k = e / c.T;
        const float h = hours[i];
        float sf, dsf;
        softfloor_eval(time[e] / h, noise[e], c, sf, dsf);
        float g = 0.f;
        if (sf > 0.f) {
            const float ot = fibre_time[k] - c.total_time;
            const float dft = 2.f * c.pfiber * (ot > 0.f ? ot : 0.01f * ot);      // d/d ot of lrelu_0.1(ot)^2
            const float d_t2 = dft - vscale * (time2[e] - class_mean[i]);
            g = gl * (class_coef[i] + h * d_t2) * dsf / h;
        }
        g_time[e] = g;
    }
}

int make_const(const pfs_loss_args& a, LossConst& c) {
    L_REQUIRE(a.S >= 2 && a.T >= 1 && (long long)a.S * a.T < (1ll << 31), "loss: bad graph sizes (at least 2 fibres)");
    c.S = a.S; c.T = a.T;
    c.total_time = a.total_time; c.wutils = a.wutils; c.wvar = a.wvar; c.pclass = a.pclass; c.pfiber = a.pfiber;
    c.noiselevel = a.noiselevel;
    c.r = a.sharpness == 0.f ? 0.f : expf(-1.f / a.sharpness);
    c.atan_r = atanf(c.r / (1.f - c.r));
    c.sharp_dev = a.sharpness_dev;
    return PFS_OK;
}
int class_row_blocks(int S) {
    int nrb = (S + 255) / 256;
    const int cap = 4 * pfs_host::sm_count();
    return nrb > cap ? cap : (nrb < 1 ? 1 : nrb);
}
int edge_grid(long long E) {
    long long b = (E + 255) / 256;
    const long long cap = 8LL * pfs_host::sm_count();
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

size_t pfs_sizeof_loss_args(void) { return sizeof(pfs_loss_args); }
size_t pfs_loss_workspace_bytes(int32_t S, int32_t T) { return (size_t)class_row_blocks(S) * 3 * T * sizeof(float) + 256; }

int pfs_loss_fwd(const pfs_loss_args* a) {
    L_REQUIRE(a && a->time && a->noise && a->hours && a->counts && a->galaxies && a->time2 && a->fibre_time && a->n_prime &&
                  a->class_mean && a->class_coef && a->scalars && a->workspace, "null pointer");
    LossConst c;
    if (int rc = make_const(*a, c)) return rc;
    cudaStream_t st = (cudaStream_t)a->stream;
    pfs_host::mark_launch(nullptr, st);
    const int nrb = class_row_blocks(c.S);
    if (a->workspace_bytes < (size_t)nrb * 3 * c.T * sizeof(float)) return lfail(PFS_ERR_WORKSPACE, "loss: workspace too small");
    float* partial = (float*)(((uintptr_t)a->workspace + 255) & ~(uintptr_t)255);
    const long long E = (long long)c.S * c.T;
    k_loss_edge_fwd<<<edge_grid(E), 256, 0, st>>>(c, a->time, a->noise, a->hours, a->galaxies, a->time2);
    L_LAUNCH_CHECK("k_loss_edge_fwd");
    k_loss_fibre_sums<<<(c.S * 32 + 255) / 256, 256, 0, st>>>(c.S, c.T, a->time2, a->fibre_time);
    L_LAUNCH_CHECK("k_loss_fibre_sums");
    const int rpb = (c.S + nrb - 1) / nrb;
    k_loss_class_partial<<<nrb, 256, 0, st>>>(c.S, c.T, a->galaxies, a->time2, rpb, partial);
    L_LAUNCH_CHECK("k_loss_class_partial");
    k_loss_scalars<<<1, 1024, 0, st>>>(c, partial, nrb, a->counts, a->time2, a->fibre_time, a->n_prime, a->class_mean,
                                       a->class_coef, a->scalars);
    L_LAUNCH_CHECK("k_loss_scalars");
    return PFS_OK;
}

int pfs_loss_bwd(const pfs_loss_args* a) {
    L_REQUIRE(a && a->time && a->noise && a->hours && a->time2 && a->fibre_time && a->class_mean && a->class_coef && a->g_time,
              "null pointer");
    LossConst c;
    if (int rc = make_const(*a, c)) return rc;
    cudaStream_t st = (cudaStream_t)a->stream;
    pfs_host::mark_launch(nullptr, st);
    k_loss_edge_bwd<<<edge_grid((long long)c.S * c.T), 256, 0, st>>>(c, a->time, a->noise, a->hours, a->time2, a->fibre_time,
                                                                    a->class_mean, a->class_coef, a->g_loss, a->g_time);
    L_LAUNCH_CHECK("k_loss_edge_bwd");
    return PFS_OK;
}

}  // extern "C"
