// Microbenchmarks behind DESIGN.md's "what bounds the kernels" notes (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_latency launch_latency.cu && ./launch_latency
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void k_empty() {}
__global__ void k_flag_exit(const unsigned char *flags, int *sink)
{
    if (flags[blockIdx.x] == 77) sink[0] = 1;
}
__global__ void k_chain(const int *next, int start, int hops, int *sink)
{
    int i = start + threadIdx.x;
    for (int h = 0; h < hops; ++h) i = next[i];
    if (i == -1) sink[0] = i;
}
// backward-like: flag -> uv -> 3 gradient loads (dependent through a predicate), no math, no stores
__global__ void k_bwd_like(const unsigned char *flags, const float2 *uv, const float *grad, int W, int H, float *sink, int mode)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + lane, py = blockIdx.y * 8 + wid, b = blockIdx.z;
    const int tilesX = W / 16, tilesY = H / 16;
    bool live = flags[(b * tilesY + (py >> 4)) * tilesX + (px >> 4)] != 0;
    if (!__any_sync(0xffffffffu, live)) return;
    const size_t pix = ((size_t)b * H + py) * W + px;
    float2 u = live ? uv[pix] : make_float2(-1.f, 0.f);
    bool c = live && u.x >= 0.f;
    if (!__any_sync(0xffffffffu, c)) return;
    if (mode == 0) { if (u.x == 123.f) sink[0] = 1.f; return; }
    const size_t plane = (size_t)W * H;
    float g0 = c ? grad[(size_t)b * 3 * plane + (size_t)py * W + px] : 0.f;
    float g1 = c ? grad[((size_t)b * 3 + 1) * plane + (size_t)py * W + px] : 0.f;
    float g2 = c ? grad[((size_t)b * 3 + 2) * plane + (size_t)py * W + px] : 0.f;
    if (g0 + g1 + g2 == 123.f) sink[0] = 1.f;
}

template <typename F> float time_us(F f, int iters = 50)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return 1e3f * ms / iters;
}

int main()
{
    int *sink; cudaMalloc(&sink, 64);
    float *fsink; cudaMalloc(&fsink, 64);
    for (int ctas : {148, 1184, 8192, 65536})
        for (int th : {32, 256})
            printf("empty kernel      %6d CTAs x %3d thr : %7.2f us\n", ctas, th, time_us([&] { k_empty<<<ctas, th>>>(); }));
    unsigned char *flags; cudaMalloc(&flags, 1 << 20); cudaMemset(flags, 0, 1 << 20);
    for (int ctas : {1184, 8192})
        printf("flag load + exit  %6d CTAs x 256 thr : %7.2f us\n", ctas, time_us([&] { k_flag_exit<<<ctas, 256>>>(flags, sink); }));
    // pointer chase: 64 MB working set (in L2 after warm-up) and 1 GB (DRAM)
    for (size_t n : {(size_t)1 << 24, (size_t)1 << 28}) {
        std::vector<int> h(n);
        const size_t stride = 4099 * 32;      // jump far, stay 128 B aligned per warp
        for (size_t i = 0; i < n; ++i) h[i] = (int)((i + stride) % n);
        int *d; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
        const int hops = 64;
        float us = time_us([&] { k_chain<<<1, 32>>>(d, 0, hops, sink); }, 20);
        printf("dependent load chain, %4zu MB array: %6.0f ns per hop (1 warp)\n", n * 4 >> 20, 1e3f * us / hops);
        cudaFree(d);
    }
    // backward-like access pattern at config 2 (8 views 512x512, 29 % of the tiles flagged, all pixels of those covered)
    const int W = 512, H = 512, B = 8;
    float2 *uv; float *grad; cudaMalloc(&uv, (size_t)B * W * H * 8); cudaMalloc(&grad, (size_t)B * 3 * W * H * 4);
    cudaMemset(uv, 0, (size_t)B * W * H * 8); cudaMemset(grad, 0, (size_t)B * 3 * W * H * 4);
    std::vector<unsigned char> hf(B * 32 * 32, 0);
    for (int b = 0; b < B; ++b) for (int y = 8; y < 25; ++y) for (int x = 8; x < 25; ++x) hf[(b * 32 + y) * 32 + x] = 1;
    cudaMemcpy(flags, hf.data(), hf.size(), cudaMemcpyHostToDevice);
    dim3 grid(W / 32, H / 8, B);
    printf("backward-like flag->uv          : %7.2f us\n", time_us([&] { k_bwd_like<<<grid, 256>>>(flags, uv, grad, W, H, fsink, 0); }));
    printf("backward-like flag->uv->3 grads : %7.2f us\n", time_us([&] { k_bwd_like<<<grid, 256>>>(flags, uv, grad, W, H, fsink, 1); }));
    cudaMemset(flags, 0, 1 << 20);
    printf("backward-like, nothing flagged  : %7.2f us\n", time_us([&] { k_bwd_like<<<grid, 256>>>(flags, uv, grad, W, H, fsink, 1); }));
    return 0;
}
