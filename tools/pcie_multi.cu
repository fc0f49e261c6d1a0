// pcie_multi.cu -- sustained host <-> device copy rate of ONE process on ONE GPU, started at a common wall-clock time so that
// several copies of it (one per GPU) measure what the box delivers to all of them together.  (measurement tool, not product)
// usage: pcie_multi <device> <h2d|d2h|both> <start unix time (s)> [seconds=3] [MB=256]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
static double now_s() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: pcie_multi <device> <h2d|d2h|both> <start> [seconds] [MB]\n"); return 1; }
    const int dev = atoi(argv[1]); const std::string mode = argv[2]; const double start = atof(argv[3]);
    const double secs = argc > 4 ? atof(argv[4]) : 3.0; const size_t n = (size_t)(argc > 5 ? atoi(argv[5]) : 256) << 20;
    cudaSetDevice(dev);
    char *h_a, *h_b, *d_a, *d_b;
    if (cudaMallocHost(&h_a, n) || cudaMallocHost(&h_b, n) || cudaMalloc(&d_a, n) || cudaMalloc(&d_b, n)) { fprintf(stderr, "alloc failed\n"); return 2; }
    memset(h_a, 1, n); memset(h_b, 2, n);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s2); cudaDeviceSynchronize();
    while (now_s() < start) { }
    const double t0 = now_s();
    double up = 0, down = 0;
    while (now_s() - t0 < secs) {
        if (mode != "d2h") { cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s1); up += n; }
        if (mode != "h2d") { cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s2); down += n; }
        cudaDeviceSynchronize();
    }
    const double t = now_s() - t0;
    printf("{\"device\": %d, \"mode\": \"%s\", \"h2d_GBps\": %.1f, \"d2h_GBps\": %.1f, \"seconds\": %.2f}\n", dev, mode.c_str(), up / t / 1e9, down / t / 1e9, t);
    return 0;
}
