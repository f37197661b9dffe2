// mtfjsp_encoder.cu -- encoder-side kernels for sm_100a: batched adjacency aggregation over the compact ELL
// adjacency the env kernel emits (no [B*N, B*N] COO matrix, no cuSPARSE).
//
// Replaces (reference file:line):
//   model/actor_critic.py:139-140, 155-156   dense adj -> to_sparse -> aggr_obs block-diagonal COO   (eliminated)
//   model/gcn_mlp.py:125                     torch.mm(Adj.double(), h.double())                       weighted neighbour sum, FP64
//   model/gcn_mlp.py:133-149                 degree via a second SpMM with ones                       fused (slot count)
//   model/gcn_mlp.py:192                     torch.sparse.mm(graph_pool, h)                           per-env mean over nodes
//
// Numerics follow the reference: products and the row sum are FP64, accumulated in ascending source index (the
// order of a coalesced COO row), divided by the in-degree including self, then rounded once to FP32.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mtfjsp.h"

namespace {

// one thread per (row, 4 channels): out[row] = (h[row] + wj*h[row-1] + wm*h[src]) / deg
__global__ void __launch_bounds__(256) aggregate_kernel(const float4* __restrict__ h, const float2* __restrict__ adj_w,
                                                        const int16_t* __restrict__ adj_src, float4* __restrict__ out,
                                                        long long rows, int N, int C4,
                                                        const float4* __restrict__ in_scale,
                                                        const float4* __restrict__ in_shift, int in_relu) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * C4) return;
    const long long row = idx / C4;
    const int c = (int)(idx - row * C4);
    const int v = (int)(row % N);
    const long long base = row - v;  // first row of this env
    const float2 w = __ldg(adj_w + row);
    const int src = __ldg(adj_src + row);
    // up to three (source, weight) pairs in ascending source order: v-1 < v always; src anywhere
    int s[3];
    double ww[3];
    int n = 0;
    const bool hj = w.x != 0.f, hm = src >= 0;
    if (hm && src < v - 1) { s[n] = src; ww[n++] = (double)w.y; }
    if (hj) { s[n] = v - 1; ww[n++] = (double)w.x; }
    if (hm && src == v - 1 && !hj) { s[n] = src; ww[n++] = (double)w.y; }
    s[n] = v; ww[n++] = 1.0;
    if (hm && src > v) { s[n] = src; ww[n++] = (double)w.y; }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_scale) { sc = __ldg(in_scale + c); sh = __ldg(in_shift + c); }
    for (int k = 0; k < n; k++) {
        float4 x = __ldg(h + (base + s[k]) * C4 + c);
        if (in_scale) {  // BatchNorm (+ReLU) of the producing layer, folded into the gather
            x.x = x.x * sc.x + sh.x; x.y = x.y * sc.y + sh.y; x.z = x.z * sc.z + sh.z; x.w = x.w * sc.w + sh.w;
            if (in_relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
        }
        a0 = fma(ww[k], (double)x.x, a0); a1 = fma(ww[k], (double)x.y, a1);
        a2 = fma(ww[k], (double)x.z, a2); a3 = fma(ww[k], (double)x.w, a3);
    }
    // weights are small integers and h is FP32, so every product is exact in FP64 and the fused form rounds the
    // same way as multiply-then-add.  Division by 1 and 2 is exact without a divide sequence.
    if (n == 3) {
        a0 = a0 / 3.0; a1 = a1 / 3.0; a2 = a2 / 3.0; a3 = a3 / 3.0;
    } else if (n == 2) {
        a0 *= 0.5; a1 *= 0.5; a2 *= 0.5; a3 *= 0.5;
    }
    out[idx] = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
}

// per-env mean over the N node rows (graph_pool average): one block per env, thread per channel
__global__ void graph_mean_kernel(const float* __restrict__ h, float* __restrict__ out, int N, int C,
                                  const float* __restrict__ in_scale, const float* __restrict__ in_shift, int in_relu) {
    const long long b = blockIdx.x;
    const float inv = 1.0f / (float)N;  // the reference multiplies each row by the float32 value 1/N and sums
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        const float sc = in_scale ? in_scale[c] : 1.f, sh = in_scale ? in_shift[c] : 0.f;
        for (int v = 0; v < N; v++) {
            float x = h[(b * N + v) * C + c];
            if (in_scale) { x = x * sc + sh; if (in_relu) x = fmaxf(x, 0.f); }
            acc += inv * x;
        }
        out[b * C + c] = acc;
    }
}

// 4-stream generalised advantage estimation (algorithm/ppo_algorithm.py:438-536): per env and reward stream a
// backward scan over the buffered steps, delta = r + gamma*v_next - v, gae = delta + gamma*lambda*gae*(1-done).
// Layout [T,B,4] (stream innermost), one thread per (env, stream): every step's load/store is coalesced.
// Also accumulates sum / sum of squares per stream (FP64) for the block-wide advantage normalisation.
__global__ void gae4_kernel(const float* __restrict__ r, const float* __restrict__ v, const float* __restrict__ vn,
                            const float* __restrict__ done, float* __restrict__ adv, double* __restrict__ stats, int T,
                            long long B, float gamma, float lam) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0, q = 0.0;
    const int k = (int)(threadIdx.x & 3);  // blockDim.x is a multiple of 4, so idx % 4 == threadIdx.x % 4
    if (idx < B * 4) {
        const long long b = idx >> 2;
        float gae = 0.f;
        for (int t = T - 1; t >= 0; t--) {
            const long long o = (long long)t * B * 4 + idx;
            const float delta = r[o] + gamma * vn[o] - v[o];
            gae = delta + gamma * lam * gae * (1.0f - done[(long long)t * B + b]);
            adv[o] = gae;
            s += (double)gae;
            q += (double)gae * (double)gae;
        }
    }
    // lanes with equal k hold the same stream: reduce over the warp with stride-4 shuffles, one atomic per warp and stream
    for (int o = 16; o >= 4; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        q += __shfl_down_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) < 4 && stats) {
        atomicAdd(stats + k, s);
        atomicAdd(stats + 4 + k, q);
    }
}

// adv <- (adv - mean) / (std + 1e-5) per stream, std unbiased (torch.std), ppo_algorithm.py:485, 532
__global__ void adv_normalize_kernel(float* __restrict__ adv, const double* __restrict__ stats, double count, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k = (int)(idx & 3);
    const double mean = stats[k] / count;
    double var = (stats[4 + k] - count * mean * mean) / (count - 1.0);
    if (var < 0) var = 0;
    adv[idx] = (float)(((double)adv[idx] - mean) / (sqrt(var) + 1e-5));
}

}  // namespace

extern "C" {

int mtfjsp_gae4(const float* r, const float* v, const float* v_next, const float* done, float* adv, double* stats, int T,
                int64_t B, float gamma, float lam, void* stream) {
    if (!r || !v || !v_next || !done || !adv || T < 1 || B < 1) return MTFJSP_E_ARG;
    const long long n = (long long)B * 4;
    gae4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r, v, v_next, done, adv, stats, T, B, gamma, lam);
    return cudaGetLastError() == cudaSuccess ? MTFJSP_OK : MTFJSP_E_CUDA;
}

int mtfjsp_adv_normalize(float* adv, const double* stats, double count, int T, int64_t B, void* stream) {
    if (!adv || !stats || count < 2 || T < 1 || B < 1) return MTFJSP_E_ARG;
    const long long total = (long long)T * B * 4;
    adv_normalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(adv, stats, count, total);
    return cudaGetLastError() == cudaSuccess ? MTFJSP_OK : MTFJSP_E_CUDA;
}


int mtfjsp_enc_aggregate(const float* h, const float* adj_w, const int16_t* adj_src, float* out, int64_t B, int N,
                         int C, const float* in_scale, const float* in_shift, int in_relu, void* stream) {
    if (!h || !adj_w || !adj_src || !out || B < 1 || N < 1 || C < 4 || (C % 4) != 0) return MTFJSP_E_ARG;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return MTFJSP_E_ARG;
    const long long rows = (long long)B * N;
    const int C4 = C / 4;
    const long long total = rows * C4;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    aggregate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(h),
                                                               reinterpret_cast<const float2*>(adj_w), adj_src,
                                                               reinterpret_cast<float4*>(out), rows, N, C4,
                                                               reinterpret_cast<const float4*>(in_scale),
                                                               reinterpret_cast<const float4*>(in_shift), in_relu);
    return cudaGetLastError() == cudaSuccess ? MTFJSP_OK : MTFJSP_E_CUDA;
}

int mtfjsp_enc_graph_mean(const float* h, float* out, int64_t B, int N, int C, const float* in_scale,
                          const float* in_shift, int in_relu, void* stream) {
    if (!h || !out || B < 1 || N < 1 || C < 1) return MTFJSP_E_ARG;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return MTFJSP_E_ARG;
    graph_mean_kernel<<<(unsigned)B, C < 128 ? 32 * ((C + 31) / 32) : 128, 0, (cudaStream_t)stream>>>(h, out, N, C, in_scale,
                                                                                                   in_shift, in_relu);
    return cudaGetLastError() == cudaSuccess ? MTFJSP_OK : MTFJSP_E_CUDA;
}

}  // extern "C"
