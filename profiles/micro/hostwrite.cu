// SM -> mapped pinned host memory write bandwidth by store pattern (what bounds the zero-copy host step).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o hostwrite hostwrite.cu && ./hostwrite
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// every warp writes `run` contiguous bytes at offset warp_id * run; W = bytes per lane store (4, 8, 16)
template <int W>
__global__ void wr(unsigned char* dst, int run, long long nwarps, int delay) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    if (delay) { long long t0 = clock64(); while (clock64() - t0 < (long long)delay * (1 + (w % 5))) {} }
    unsigned char* p = dst + w * run;
    for (int o = lane * W; o < run; o += 32 * W) {
        if (W == 16) *reinterpret_cast<uint4*>(p + o) = make_uint4(w, o, 1, 2);
        else if (W == 8) *reinterpret_cast<uint2*>(p + o) = make_uint2(w, o);
        else *reinterpret_cast<uint32_t*>(p + o) = (uint32_t)w;
    }
}
// copy a device buffer to host with SM loads/stores, 16 B per lane, fully coalesced
__global__ void cp(const uint4* src, uint4* dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main() {
    const size_t total = 4718592 * 4;  // 4 x the J6M6 x 65,536 record volume
    unsigned char *h, *d;
    cudaHostAlloc(&h, total, cudaHostAllocDefault);
    cudaMalloc(&d, total);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, auto launch, size_t bytes) {
        float best = 1e9f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-64s %8.1f us  %6.1f GB/s  (%s)\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    time("cudaMemcpyAsync D2H 4.7 MB", [&] { cudaMemcpyAsync(h, d, 4718592, cudaMemcpyDeviceToHost); }, 4718592);
    time("cudaMemcpyAsync D2H 18.9 MB", [&] { cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost); }, total);
    for (size_t tot : {(size_t)4718592, total}) {
        printf("-- total %zu bytes\n", tot);
        time("SM copy kernel 16 B/lane, 148x8 blocks x 256", [&] { cp<<<148 * 8, 256>>>((const uint4*)d, (uint4*)h, tot / 16); }, tot);
        time("SM copy kernel 16 B/lane, 148x2 blocks x 256", [&] { cp<<<148 * 2, 256>>>((const uint4*)d, (uint4*)h, tot / 16); }, tot);
        for (int run : {288, 256, 512, 1024, 4096}) {
            const long long nw = tot / run;
            char nm[96];
            snprintf(nm, sizeof nm, "warp runs of %d B, 16 B/lane", run);
            if (run % 16 == 0) time(nm, [&] { wr<16><<<(unsigned)((nw * 32 + 127) / 128), 128>>>(h, run, nw, 0); }, nw * run);
            snprintf(nm, sizeof nm, "warp runs of %d B, 8 B/lane", run);
            time(nm, [&] { wr<8><<<(unsigned)((nw * 32 + 127) / 128), 128>>>(h, run, nw, 0); }, nw * run);
            snprintf(nm, sizeof nm, "warp runs of %d B, 4 B/lane", run);
            time(nm, [&] { wr<4><<<(unsigned)((nw * 32 + 127) / 128), 128>>>(h, run, nw, 0); }, nw * run);
        }
        {
            const long long nw = tot / 288;
            time("warp runs of 288 B, 8 B/lane, staggered by ~2-10 us spins", [&] { wr<8><<<(unsigned)((nw * 32 + 127) / 128), 128>>>(h, 288, nw, 4000); }, nw * 288);
        }
    }
    return 0;
}
