// Library-wide state of libb200audio: error text, launch counter, device check.
#include "common.cuh"

#include <cstdlib>

namespace b2a {

static thread_local std::string t_last_error;
std::atomic<long long> g_launches{0};

void set_last_error(const std::string& msg) { t_last_error = msg; }

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B2A_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        throw Error(B2A_ERR_CUDA, "no CUDA device visible: libb200audio has no CPU fallback");
    }
    if (device < 0 || device >= n) throw Error(B2A_ERR_CUDA, "CUDA device index out of range");
    B2A_CUDA(cudaSetDevice(device));
    cudaDeviceProp p{};
    B2A_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10)
        throw Error(B2A_ERR_CUDA, std::string("libb200audio is built for sm_100a only; found ") + p.name);
}

}  // namespace b2a

extern "C" {

const char* b2a_last_error(void) { return b2a::t_last_error.c_str(); }
const char* b2a_version(void) { return "b200audio 0.1 (sm_100a)"; }

int32_t b2a_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t b2a_launch_count(void) { return b2a::g_launches.load(); }

}  // extern "C"
