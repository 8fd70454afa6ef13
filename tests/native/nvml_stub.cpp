// Test-only NVML: a shared library with the entry points csrc/nvml_dyn.hpp resolves, describing a synthetic node so
// that the NVML-driven enumeration (MIG devices, probe=off) can be exercised without hardware.  Loaded through
// B2DP_NVML_LIBRARY; never part of the product.
//
//   B2DP_NVML_STUB="gpus=8,mig=7"            8 GPUs, MIG enabled on all, 7 x 1g.23gb instances each
//   B2DP_NVML_STUB="gpus=4,mig=3,migmask=5"  MIG enabled on GPUs 0 and 2 only (3 instances each), 1 and 3 whole
//   B2DP_NVML_STUB="gpus=2,mig=0"            two whole GPUs; minors are reversed (minor != NVML index)
//   B2DP_NVML_STUB="gpus=2,mig=3,empty=2"    GPU 1 has MIG mode enabled but no instance has been created on it
// PCI layout = k8s-device-plugin_b200/synth.py:write_b200_tree: bus 0x19 + 0x10*g, device 0, domain 0.
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
struct Dev { int gpu; int slot; };  // slot < 0: the physical GPU
Dev g_gpu[16];
Dev g_mig[16][8];
int g_n = 8, g_mig_n = 7, g_mask = -1, g_empty = 0;  // g_empty: GPUs with MIG mode enabled but no instance created
bool g_init = false;

void configure() {
    if (g_init) return;
    g_init = true;
    const char* c = getenv("B2DP_NVML_STUB");
    if (c) {
        const char* p;
        if ((p = strstr(c, "gpus="))) g_n = atoi(p + 5);
        if ((p = strstr(c, "mig="))) g_mig_n = atoi(p + 4);
        if ((p = strstr(c, "migmask="))) g_mask = atoi(p + 8);
        if ((p = strstr(c, "empty="))) g_empty = atoi(p + 6);
    }
    if (g_n < 1) g_n = 1;
    if (g_n > 16) g_n = 16;
    if (g_mig_n > 7) g_mig_n = 7;
    for (int g = 0; g < 16; ++g) {
        g_gpu[g] = {g, -1};
        for (int s = 0; s < 8; ++s) g_mig[g][s] = {g, s};
    }
}
bool mig_on(int g) { return g_mig_n > 0 && (g_mask < 0 || (g_mask >> g) & 1); }
const Dev* dev(void* h) { return static_cast<const Dev*>(h); }
int bus_of(int g) { return 0x19 + 0x10 * g; }
void put(char* dst, unsigned len, const char* s) { snprintf(dst, len, "%s", s); }
}  // namespace

struct PciInfo { char busIdLegacy[16]; unsigned domain, bus, device, pciDeviceId, pciSubSystemId; char busId[32]; };
struct Attributes { unsigned mp, ce, dec, enc, jpg, ofa, giSlices, ciSlices; unsigned long long memMB; };
struct Memory { unsigned long long total, free_, used; };

extern "C" {
#define API __attribute__((visibility("default")))
API int nvmlInit_v2() { configure(); return 0; }
API int nvmlShutdown() { return 0; }
API int nvmlSystemGetDriverVersion(char* v, unsigned len) { put(v, len, "580.159.03"); return 0; }
API int nvmlDeviceGetCount_v2(unsigned* n) { configure(); *n = (unsigned)g_n; return 0; }
API int nvmlDeviceGetHandleByIndex_v2(unsigned i, void** h) { configure(); if ((int)i >= g_n) return 2; *h = &g_gpu[i]; return 0; }
API int nvmlDeviceGetHandleByPciBusId_v2(const char* id, void** h) {
    configure();
    unsigned dom = 0, bus = 0, d = 0;
    if (sscanf(id, "%x:%x:%x", &dom, &bus, &d) != 3) return 2;
    for (int g = 0; g < g_n; ++g) if ((unsigned)bus_of(g) == bus) { *h = &g_gpu[g]; return 0; }
    return 6;
}
API int nvmlDeviceGetPciInfo_v3(void* h, PciInfo* pi) {
    memset(pi, 0, sizeof *pi);
    pi->domain = 0; pi->bus = (unsigned)bus_of(dev(h)->gpu); pi->device = 0; pi->pciDeviceId = 0x290110de;
    snprintf(pi->busId, sizeof pi->busId, "%08X:%02X:%02X.0", 0, bus_of(dev(h)->gpu), 0);
    snprintf(pi->busIdLegacy, sizeof pi->busIdLegacy, "%04X:%02X:%02X.0", 0, bus_of(dev(h)->gpu), 0);
    return 0;
}
API int nvmlDeviceGetMinorNumber(void* h, unsigned* m) { *m = (unsigned)(g_n - 1 - dev(h)->gpu); return 0; }  // minor != index on purpose
API int nvmlDeviceGetIndex(void* h, unsigned* i) { *i = (unsigned)dev(h)->gpu; return 0; }
API int nvmlDeviceGetUUID(void* h, char* u, unsigned len) {
    if (dev(h)->slot < 0) snprintf(u, len, "GPU-%08x-0000-4000-8000-00000000b200", dev(h)->gpu);
    else snprintf(u, len, "MIG-%08x-%04x-4000-8000-00000000b200", dev(h)->gpu, dev(h)->slot);
    return 0;
}
API int nvmlDeviceGetName(void* h, char* n, unsigned len) { put(n, len, dev(h)->slot < 0 ? "NVIDIA B200" : "NVIDIA B200 MIG 1g.23gb"); return 0; }
API int nvmlDeviceGetVbiosVersion(void*, char* v, unsigned len) { put(v, len, "97.00.88.00.0F"); return 0; }
API int nvmlDeviceGetInforomImageVersion(void*, char* v, unsigned len) { put(v, len, "G548.0201.00.06"); return 0; }
API int nvmlDeviceGetInforomVersion(void*, int obj, char* v, unsigned len) { put(v, len, obj == 0 ? "2.1" : obj == 1 ? "7.16" : "N/A x"); return obj > 2 ? 3 : 0; }
API int nvmlDeviceGetGspFirmwareVersion(void*, char* v) { put(v, 64, "580.159.03"); return 0; }
API int nvmlDeviceGetMigMode(void* h, unsigned* cur, unsigned* pend) { *cur = *pend = mig_on(dev(h)->gpu) ? 1u : 0u; return 0; }
API int nvmlDeviceGetMaxMigDeviceCount(void*, unsigned* n) { *n = 7; return 0; }
API int nvmlDeviceGetMigDeviceHandleByIndex(void* h, unsigned idx, void** mh) {
    if (!mig_on(dev(h)->gpu) || (int)idx >= g_mig_n || ((g_empty >> dev(h)->gpu) & 1)) return 6;  // NVML_ERROR_NOT_FOUND
    *mh = &g_mig[dev(h)->gpu][idx];
    return 0;
}
API int nvmlDeviceGetGpuInstanceId(void* h, unsigned* id) { *id = 7u + (unsigned)dev(h)->slot; return 0; }
API int nvmlDeviceGetComputeInstanceId(void*, unsigned* id) { *id = 0; return 0; }
API int nvmlDeviceGetDeviceHandleFromMigDeviceHandle(void* mh, void** h) { *h = &g_gpu[dev(mh)->gpu]; return 0; }
API int nvmlDeviceGetAttributes_v2(void* h, Attributes* a) {
    if (dev(h)->slot < 0) return 3;  // NOT_SUPPORTED on a physical GPU
    memset(a, 0, sizeof *a);
    a->mp = 18; a->giSlices = 1; a->ciSlices = 1; a->memMB = 23552;
    return 0;
}
API int nvmlDeviceGetMemoryInfo(void*, Memory* m) { m->total = 192265846784ull; m->free_ = m->total; m->used = 0; return 0; }
API int nvmlDeviceGetNumGpuCores(void*, unsigned* c) { *c = 148 * 128; return 0; }
API int nvmlDeviceGetCudaComputeCapability(void*, int* M, int* m) { *M = 10; *m = 0; return 0; }
API int nvmlDeviceGetNvLinkState(void*, unsigned, unsigned* active) { *active = 1; return 0; }
API int nvmlDeviceGetP2PStatus(void* a, void* b, int, int* status) { *status = (mig_on(dev(a)->gpu) || mig_on(dev(b)->gpu)) ? 3 : 0; return 0; }
API int nvmlDeviceGetComputeRunningProcesses_v3(void*, unsigned* n, void*) { *n = 0; return 0; }
API int nvmlDeviceGetTotalEccErrors(void*, int, int, unsigned long long* c) { *c = 0; return 0; }
}
