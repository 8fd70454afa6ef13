// p2p_sweep.cu — tuning harness for the NVLink P2P probe (needs 2 GPUs with peer access).
// Reader GPU pulls a peer-mapped buffer into local HBM with the probe kernels of hbm_probe.cuh
// (verify + checksum included); reports GB/s of bytes crossing NVLink per direction, for one
// direction alone and for both directions at once, next to cudaMemcpyPeerAsync.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
//        -I k8s-device-plugin_b200/csrc tools/p2p_sweep.cu -o tools/p2p_sweep
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "hbm_probe.cuh"

using namespace b2dp;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

struct Dev {
    int id; uint4 *pat, *dst; ProbeCtl* ctl; ProbeOut *out_h, *out_d; cudaStream_t st; cudaEvent_t e0, e1; int sms;
    uint32_t seed;
};
using LaunchFn = void (*)(Dev& d, const uint4* src, unsigned long long n_vec, uint32_t seed, int ctas_per_sm, unsigned long long seq);

template <int T, int U> static void l_r128(Dev& d, const uint4* src, unsigned long long n, uint32_t seed, int k, unsigned long long seq) {
    hbm_probe_r128<T, U><<<d.sms * k, T, 0, d.st>>>(src, d.dst, n, seed, 0u, d.ctl, d.out_d, seq);
}
template <int T, int U> static void l_r256(Dev& d, const uint4* src, unsigned long long n, uint32_t seed, int k, unsigned long long seq) {
    hbm_probe_r256<T, U><<<d.sms * k, T, 0, d.st>>>(src, d.dst, n, seed, 0u, d.ctl, d.out_d, seq);
}
template <int CW, int TV, int ST> static void l_tma(Dev& d, const uint4* src, unsigned long long n, uint32_t seed, int k, unsigned long long seq) {
    constexpr size_t smem = (size_t)ST * TV * 16 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(hbm_probe_tma<CW, TV, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hbm_probe_tma<CW, TV, ST><<<d.sms * k, (CW + 1) * 32, smem, d.st>>>(src, d.dst, n, seed, 0u, d.ctl, d.out_d, seq);
}
struct Cfg { std::string name; LaunchFn fn; int k; };

int main(int argc, char** argv) {
    unsigned long long bytes = 256ull << 20;
    int iters = 6;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--bytes")) bytes = strtoull(argv[++i], 0, 0);
        else if (!strcmp(argv[i], "--iters")) iters = atoi(argv[++i]);
    }
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { fprintf(stderr, "need 2 GPUs\n"); return 1; }
    const unsigned long long n_vec = bytes / 16;
    Dev D[2];
    for (int i = 0; i < 2; ++i) {
        Dev& d = D[i]; d.id = i; d.seed = 0x5EED0000u | i;
        CK(cudaSetDevice(i));
        int can = 0; CK(cudaDeviceCanAccessPeer(&can, i, 1 - i));
        if (!can) { fprintf(stderr, "no peer access %d->%d\n", i, 1 - i); return 1; }
        CK(cudaDeviceEnablePeerAccess(1 - i, 0));
        cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, i)); d.sms = p.multiProcessorCount;
        CK(cudaMalloc(&d.pat, bytes)); CK(cudaMalloc(&d.dst, bytes)); CK(cudaMalloc(&d.ctl, sizeof(ProbeCtl)));
        ProbeCtl init{}; init.first_bad = ~0ull; init.t_start_ns = ~0ull;
        CK(cudaMemcpy(d.ctl, &init, sizeof init, cudaMemcpyHostToDevice));
        CK(cudaHostAlloc(&d.out_h, sizeof(ProbeOut), cudaHostAllocMapped | cudaHostAllocPortable));
        CK(cudaHostGetDevicePointer(&d.out_d, d.out_h, 0));
        CK(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&d.e0)); CK(cudaEventCreate(&d.e1));
        hbm_fill<256><<<d.sms * 8, 256, 0, d.st>>>(d.pat, n_vec, d.seed);
        CK(cudaStreamSynchronize(d.st));
    }
    std::vector<Cfg> cfgs;
#define R128(T, U) for (int k : {1, 2, 4, 8}) if (k * T <= 2048) cfgs.push_back({"r128_t" #T "_u" #U, l_r128<T, U>, k});
#define R256(T, U) for (int k : {1, 2, 4}) if (k * T <= 2048) cfgs.push_back({"r256_t" #T "_u" #U, l_r256<T, U>, k});
#define TMA(CW, TV, ST) for (int k : {1, 2, 3, 4}) if ((size_t)k * ((size_t)ST * TV * 16 + 1024) <= 227 * 1024) cfgs.push_back({"tma_cw" #CW "_tv" #TV "_st" #ST, l_tma<CW, TV, ST>, k});
    R128(256, 4) R128(256, 8) R128(512, 2) R128(512, 4) R128(512, 8) R128(1024, 2) R128(1024, 4)
    R256(256, 2) R256(256, 4) R256(512, 2) R256(512, 4)
    TMA(4, 1024, 3) TMA(4, 1024, 4) TMA(4, 1024, 6) TMA(4, 2048, 3) TMA(4, 2048, 4) TMA(4, 512, 6) TMA(8, 2048, 4) TMA(4, 4096, 3)

    printf("name,ctas_per_sm,uni_gbs,bi_gbs_per_dir,ok\n");
    // baseline: driver peer copy
    {
        float best_u = 1e30f, best_b = 1e30f;
        for (int it = 0; it < iters + 1; ++it) {
            CK(cudaSetDevice(0));
            CK(cudaEventRecord(D[0].e0, D[0].st));
            CK(cudaMemcpyPeerAsync(D[0].dst, 0, D[1].pat, 1, bytes, D[0].st));
            CK(cudaEventRecord(D[0].e1, D[0].st)); CK(cudaEventSynchronize(D[0].e1));
            float ms; CK(cudaEventElapsedTime(&ms, D[0].e0, D[0].e1)); if (it) best_u = std::min(best_u, ms);
        }
        for (int it = 0; it < iters + 1; ++it) {
            for (int i = 0; i < 2; ++i) { CK(cudaSetDevice(i)); CK(cudaEventRecord(D[i].e0, D[i].st));
                CK(cudaMemcpyPeerAsync(D[i].dst, i, D[1 - i].pat, 1 - i, bytes, D[i].st)); CK(cudaEventRecord(D[i].e1, D[i].st)); }
            float worst = 0;
            for (int i = 0; i < 2; ++i) { CK(cudaEventSynchronize(D[i].e1)); float ms; CK(cudaEventElapsedTime(&ms, D[i].e0, D[i].e1)); worst = std::max(worst, ms); }
            if (it) best_b = std::min(best_b, worst);
        }
        printf("cudaMemcpyPeerAsync,0,%.1f,%.1f,1\n", bytes / best_u / 1e6, bytes / best_b / 1e6);
    }
    unsigned long long seq = 0;
    for (auto& c : cfgs) {
        float best_u = 1e30f, best_b = 1e30f; bool ok = true;
        for (int it = 0; it < iters + 1; ++it) {
            CK(cudaSetDevice(0));
            CK(cudaEventRecord(D[0].e0, D[0].st));
            c.fn(D[0], D[1].pat, n_vec, D[1].seed, c.k, ++seq);
            CK(cudaGetLastError());
            CK(cudaEventRecord(D[0].e1, D[0].st)); CK(cudaEventSynchronize(D[0].e1));
            float ms; CK(cudaEventElapsedTime(&ms, D[0].e0, D[0].e1)); if (it) best_u = std::min(best_u, ms);
            if (D[0].out_h->mismatches != 0 || D[0].out_h->seq != seq) ok = false;
        }
        for (int it = 0; it < iters + 1; ++it) {
            unsigned long long s[2];
            for (int i = 0; i < 2; ++i) { CK(cudaSetDevice(i)); CK(cudaEventRecord(D[i].e0, D[i].st));
                c.fn(D[i], D[1 - i].pat, n_vec, D[1 - i].seed, c.k, s[i] = ++seq); CK(cudaGetLastError()); CK(cudaEventRecord(D[i].e1, D[i].st)); }
            float worst = 0;
            for (int i = 0; i < 2; ++i) { CK(cudaEventSynchronize(D[i].e1)); float ms; CK(cudaEventElapsedTime(&ms, D[i].e0, D[i].e1)); worst = std::max(worst, ms);
                if (D[i].out_h->mismatches != 0 || D[i].out_h->seq != s[i]) ok = false; }
            if (it) best_b = std::min(best_b, worst);
        }
        printf("%s,%d,%.1f,%.1f,%d\n", c.name.c_str(), c.k, bytes / best_u / 1e6, bytes / best_b / 1e6, ok ? 1 : 0);
        fflush(stdout);
    }
    return 0;
}
