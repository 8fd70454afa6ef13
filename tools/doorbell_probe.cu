// doorbell_probe.cu -- feasibility measurement: how long from "the host wants the pass to start" to "the host sees the
// kernel's published flag", for (a) an ordinary launch and (b) a kernel enqueued in advance behind
// cuStreamWaitValue32 on a host-mapped doorbell and released by a plain host store.  With idle gaps before the release
// (a heartbeat comes every few seconds in production), and for N GPUs released back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/doorbell_probe tools/doorbell_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

__global__ void publish(volatile unsigned long long* out, unsigned long long seq) {
    __threadfence_system();
    *out = seq;
}

static double now_us() {
    using namespace std::chrono;
    return duration<double, std::micro>(steady_clock::now().time_since_epoch()).count();
}
static double med(std::vector<double> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }
static double p99(std::vector<double> v) { std::sort(v.begin(), v.end()); return v[(size_t)(0.99 * (v.size() - 1))]; }

struct G {
    int dev;
    cudaStream_t s;
    volatile unsigned long long* out_h;
    unsigned long long* out_d;
    volatile unsigned int* bell_h;
    CUdeviceptr bell_d;
};

int main(int argc, char** argv) {
    int n = 0;
    cudaGetDeviceCount(&n);
    if (argc > 1) n = std::min(n, atoi(argv[1]));
    std::vector<G> g(n);
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(i);
        g[i].dev = i;
        cudaStreamCreateWithFlags(&g[i].s, cudaStreamNonBlocking);
        void* p;
        cudaHostAlloc(&p, 64, cudaHostAllocMapped);
        g[i].out_h = (volatile unsigned long long*)p;
        *g[i].out_h = 0;
        cudaHostGetDevicePointer((void**)&g[i].out_d, p, 0);
        cudaHostAlloc(&p, 64, cudaHostAllocMapped);
        g[i].bell_h = (volatile unsigned int*)p;
        *g[i].bell_h = 0;
        void* dp;
        cudaHostGetDevicePointer(&dp, p, 0);
        g[i].bell_d = (CUdeviceptr)dp;
        publish<<<1, 1, 0, g[i].s>>>(g[i].out_d, 0);
        cudaStreamSynchronize(g[i].s);
    }
    unsigned long long seq = 0;
    printf("mode,n_gpus,idle_ms,median_us,p99_us,max_us\n");
    for (int idle_ms : {0, 1, 100, 2000}) {
        const int reps = idle_ms >= 2000 ? 8 : idle_ms >= 100 ? 30 : 300;
        for (int mode = 0; mode < 2; ++mode) {
            std::vector<double> lat;
            for (int r = 0; r < reps; ++r) {
                ++seq;
                if (mode == 1)  // arm: wait for the doorbell, then the kernel
                    for (auto& x : g) {
                        cudaSetDevice(x.dev);
                        cuStreamWaitValue32((CUstream)x.s, x.bell_d, (unsigned)seq, CU_STREAM_WAIT_VALUE_EQ);
                        publish<<<1, 1, 0, x.s>>>(x.out_d, seq);
                    }
                if (idle_ms) std::this_thread::sleep_for(std::chrono::milliseconds(idle_ms));
                const double t0 = now_us();
                if (mode == 0)
                    for (auto& x : g) { cudaSetDevice(x.dev); publish<<<1, 1, 0, x.s>>>(x.out_d, seq); }
                else
                    for (auto& x : g) *x.bell_h = (unsigned)seq;
                for (auto& x : g) while (*x.out_h != seq) {}
                lat.push_back(now_us() - t0);
            }
            printf("%s,%d,%d,%.2f,%.2f,%.2f\n", mode ? "doorbell" : "launch", n, idle_ms, med(lat), p99(lat), *std::max_element(lat.begin(), lat.end()));
        }
    }
    return 0;
}
