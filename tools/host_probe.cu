// host_probe.cu -- what the host side of a GPU box can do: PCIe H2D / D2H / both at once (pinned memory),
// and multi-threaded u8 -> u16 widening (the cost of a narrow D2H transport for the host entry points).
// Build: nvcc -O2 -Xcompiler -mavx2,-pthread tools/host_probe.cu -o tools/_build/host_probe
#include <cuda_runtime.h>
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void widen(const uint8_t* s, uint16_t* d, size_t n, bool nt) {
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256i v = _mm256_loadu_si256((const __m256i*)(s + i));
        __m256i lo = _mm256_cvtepu8_epi16(_mm256_castsi256_si128(v));
        __m256i hi = _mm256_cvtepu8_epi16(_mm256_extracti128_si256(v, 1));
        if (nt) { _mm256_stream_si256((__m256i*)(d + i), lo); _mm256_stream_si256((__m256i*)(d + i + 16), hi); }
        else { _mm256_storeu_si256((__m256i*)(d + i), lo); _mm256_storeu_si256((__m256i*)(d + i + 16), hi); }
    }
    for (; i < n; i++) d[i] = s[i];
}
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
    const size_t n = 150u << 20;
    printf("{\"hw_threads\": %u}\n", std::thread::hardware_concurrency());
    uint8_t *h_in, *h_out8; uint16_t* h_out; uint8_t *d_a, *d_b;
    cudaMallocHost(&h_in, n); cudaMallocHost(&h_out, n * 2); cudaMallocHost(&h_out8, n);
    cudaMalloc(&d_a, n); cudaMalloc(&d_b, n * 2);
    memset(h_in, 1, n); memset(h_out, 0, n * 2); memset(h_out8, 2, n);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto fn, double bytes) {
        float best = 1e9;
        for (int r = 0; r < 5; r++) {
            cudaDeviceSynchronize();
            double t0 = now_ms(); fn(); cudaDeviceSynchronize(); double ms = now_ms() - t0;
            if (ms < best) best = (float)ms;
        }
        printf("{\"test\": \"%s\", \"ms\": %.3f, \"GBps\": %.1f}\n", name, best, bytes / best / 1e6);
    };
    timeit("h2d_150MB", [&] { cudaMemcpyAsync(d_a, h_in, n, cudaMemcpyHostToDevice, s1); }, n);
    timeit("d2h_300MB", [&] { cudaMemcpyAsync(h_out, d_b, 2 * n, cudaMemcpyDeviceToHost, s2); }, 2.0 * n);
    timeit("bidir_150up_300down", [&] { cudaMemcpyAsync(d_a, h_in, n, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h_out, d_b, 2 * n, cudaMemcpyDeviceToHost, s2); }, 3.0 * n);
    timeit("bidir_150up_150down", [&] { cudaMemcpyAsync(d_a, h_in, n, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h_out8, d_b, n, cudaMemcpyDeviceToHost, s2); }, 2.0 * n);
    timeit("bidir_sliced16MB_150up_300down", [&] {
        for (size_t o = 0; o < n; o += 16u << 20) { size_t m = std::min<size_t>(16u << 20, n - o);
            cudaMemcpyAsync(d_a + o, h_in + o, m, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync((uint8_t*)h_out + 2 * o, d_b + 2 * o, 2 * m, cudaMemcpyDeviceToHost, s2); } }, 3.0 * n);
    for (int nt = 0; nt < 2; nt++)
        for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
            if (T > (int)std::thread::hardware_concurrency()) break;
            double best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                double t0 = now_ms();
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++) th.emplace_back([=] { size_t a = n / T * t / 32 * 32, b = t == T - 1 ? n : n / T * (t + 1) / 32 * 32; widen(h_out8 + a, h_out + a, b - a, nt); });
                for (auto& x : th) x.join();
                best = std::min(best, now_ms() - t0);
            }
            printf("{\"test\": \"widen_u8_u16\", \"nt_stores\": %d, \"threads\": %d, \"ms\": %.2f, \"out_GBps\": %.1f}\n", nt, T, best, n * 2 / best / 1e6);
        }
    // widening while both DMA directions are busy
    for (int T : {4, 8, 16}) {
        if (T > (int)std::thread::hardware_concurrency()) break;
        cudaDeviceSynchronize();
        double t0 = now_ms();
        cudaMemcpyAsync(d_a, h_in, n, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h_out8, d_b, n, cudaMemcpyDeviceToHost, s2);
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([=] { size_t a = n / T * t / 32 * 32, b = t == T - 1 ? n : n / T * (t + 1) / 32 * 32; widen(h_in + a, h_out + a, b - a, true); });
        for (auto& x : th) x.join();
        double tw = now_ms() - t0;
        cudaDeviceSynchronize();
        printf("{\"test\": \"widen_during_bidir_dma\", \"threads\": %d, \"widen_ms\": %.2f, \"all_ms\": %.2f}\n", T, tw, now_ms() - t0);
    }
    return 0;
}
