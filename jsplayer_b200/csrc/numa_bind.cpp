// numa_bind.cpp -- keep a GPU's host thread and its pinned staging memory on the GPU's NUMA node ("scaling is
// limited by host PCIe/NUMA placement, so pin host threads and staging memory to the GPU's NUMA node").
//
// The end-to-end path moves 4 bytes per decoded pixel over PCIe (the Int32Array contract of IVideoCodec.hx:11-29).  With one
// rank per GPU on a two-socket box, a rank whose pinned pictures live on the other socket pays the inter-socket link on every
// D2H write.  Everything here is plain Linux: sysfs for the topology, sched_setaffinity(2) and set_mempolicy(2) by syscall number
// (no libnuma in the image).  Disabled with JSP_NUMA_BIND=0; a box with a single node (or a hidden topology) is left alone.
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>
#include <sys/syscall.h>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace jsp {

namespace {
constexpr int MPOL_DEFAULT_ = 0, MPOL_PREFERRED_ = 1;

bool read_line(const std::string &path, std::string &out)
{
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096];
    const bool ok = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!ok) return false;
    out = buf;
    while (!out.empty() && isspace((unsigned char)out.back())) out.pop_back();
    return true;
}

// "0-15,32-47" -> cpu ids
std::vector<int> parse_cpulist(const std::string &s)
{
    std::vector<int> v;
    size_t i = 0;
    while (i < s.size()) {
        char *end = nullptr;
        const long a = strtol(s.c_str() + i, &end, 10);
        if (end == s.c_str() + i) break;
        long b = a;
        i = (size_t)(end - s.c_str());
        if (i < s.size() && s[i] == '-') {
            b = strtol(s.c_str() + i + 1, &end, 10);
            i = (size_t)(end - s.c_str());
        }
        for (long c = a; c <= b && c < 4096; c++) v.push_back((int)c);
        if (i < s.size() && s[i] == ',') i++;
    }
    return v;
}

int count_nodes()
{
    std::string s;
    if (!read_line("/sys/devices/system/node/online", s)) return 1;
    return (int)parse_cpulist(s).size();
}

bool enabled()
{
    const char *e = getenv("JSP_NUMA_BIND");
    return !(e && e[0] == '0');
}
}  // namespace

// NUMA node of a CUDA device from sysfs, or -1 (unknown / single node / virtualised without topology).
int device_numa_node(int device)
{
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *p = bus; *p; ++p) *p = (char)tolower((unsigned char)*p);
    std::string s;
    if (!read_line(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", s)) return -1;
    const int node = atoi(s.c_str());
    return node >= 0 ? node : -1;
}

// Binds the CALLING thread to the CPUs of `device`'s NUMA node (intersected with the CPUs it may already use) and makes that
// node the preferred one for its page allocations -- pinned memory allocated from this thread afterwards is node-local.
// Returns the node, or -1 when nothing was changed.
int bind_thread_to_device(int device)
{
    if (!enabled() || count_nodes() < 2) return -1;
    const int node = device_numa_node(device);
    if (node < 0) return -1;
    std::string s;
    if (!read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", s)) return -1;
    const std::vector<int> cpus = parse_cpulist(s);
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed); CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return -1;
    int n = 0;
    for (int c : cpus) if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); n++; }
    if (n == 0) return -1;                                           // the container does not own that node's CPUs: leave it
    if (sched_setaffinity(0, sizeof want, &want) != 0) return -1;
    if (node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_set_mempolicy, MPOL_PREFERRED_, mask, (unsigned long)(8 * sizeof mask + 1));   // failure = keep the default policy
    }
    return node;
}

}  // namespace jsp

// What this host takes from `device` over PCIe: `reps` device -> pinned-host copies of `bytes` each, GB/s (0 on failure).  Ranks
// that call it at the same time (after a barrier) measure the box's ceiling for N concurrent streams -- the bound of every
// end-to-end number of this library (4 bytes per decoded pixel), and different from box to box.
static double host_d2h_gbs(int device, size_t bytes, int reps)
{
    if (bytes == 0 || reps <= 0 || cudaSetDevice(device) != cudaSuccess) return 0.0;
    void *d = nullptr, *h = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double gbs = 0.0;
    if (cudaMalloc(&d, bytes) == cudaSuccess && cudaHostAlloc(&h, bytes, cudaHostAllocDefault) == cudaSuccess &&
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess &&
        cudaEventCreate(&e1) == cudaSuccess) {
        cudaMemsetAsync(d, 1, bytes, st);
        cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st);      // touches every page once
        cudaEventRecord(e0, st);
        for (int i = 0; i < reps; i++) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st);
        cudaEventRecord(e1, st);
        float ms = 0.f;
        if (cudaStreamSynchronize(st) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0.f)
            gbs = (double)bytes * reps / (ms * 1e-3) / 1e9;
    }
    cudaGetLastError();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    return gbs;
}

extern "C" {
__attribute__((visibility("default"))) double jsp_host_d2h_gbs(int device, size_t bytes, int reps) { return host_d2h_gbs(device, bytes, reps); }
__attribute__((visibility("default"))) int jsp_numa_node_of_device(int device) { return jsp::device_numa_node(device); }
__attribute__((visibility("default"))) int jsp_numa_bind_thread(int device) { return jsp::bind_thread_to_device(device); }
}
