// Error plumbing and library-level entry points of libipdm_b200.so (include/ipdm_b200.h).
#include "common.cuh"
#include <algorithm>

#include <cstring>
#include <vector>

namespace ipdm {

static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

bool g_prof_on = false;
struct ProfRec { int kind; cudaEvent_t a, b; double work, bytes; };
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_prof_open[PROF_KINDS];
void prof_begin(int kind, cudaStream_t st) {
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); g_prof_open[kind] = e;
}
void prof_end(int kind, cudaStream_t st, double work, double bytes) {
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
    g_prof.push_back({kind, g_prof_open[kind], e, work, bytes});
}

}  // namespace ipdm

extern "C" void ipdm_profile_enable(int on) {
    ipdm::g_prof_on = on != 0;
    if (on) { for (auto& r : ipdm::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } ipdm::g_prof.clear(); }
}
extern "C" int ipdm_profile_collect(double* ms_out, double* work_out, long long* launches_out) {
    using namespace ipdm;
    IPDM_CHECK_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < PROF_KINDS; ++k) { ms_out[k] = 0; work_out[k] = 0; launches_out[k] = 0; }
    for (auto& r : g_prof) {
        float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
        ms_out[r.kind] += ms; work_out[r.kind] += r.work; launches_out[r.kind] += 1;
    }
    return IPDM_OK;
}

// Roofline of a FLOP-counted family whose launches are not all on the same side of the ridge (the halo conv family: 128/256-channel
// layers are tensor-bound, the 64-channel image layers HBM-bound with an fp32 residual stream): sum over launches of
// max(flops / peak_flops, algorithmic bytes / peak_bytes) = the time the family would take with every launch AT its own roof.
extern "C" int ipdm_profile_roofline(int kind, double peak_flops_per_s, double peak_bytes_per_s, double* roof_ms_out, double* ms_out,
                                     double* bytes_out) {
    using namespace ipdm;
    IPDM_REQUIRE(kind >= 0 && kind < PROF_KINDS && peak_flops_per_s > 0 && peak_bytes_per_s > 0 && roof_ms_out && ms_out, "ipdm_profile_roofline: bad arguments");
    IPDM_CHECK_CUDA(cudaDeviceSynchronize());
    double roof = 0, tot = 0, by = 0;
    for (auto& r : g_prof) {
        if (r.kind != kind) continue;
        float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
        tot += ms; by += r.bytes;
        roof += 1e3 * std::max(r.work / peak_flops_per_s, r.bytes / peak_bytes_per_s);
    }
    *roof_ms_out = roof; *ms_out = tot; if (bytes_out) *bytes_out = by;
    return IPDM_OK;
}

extern "C" const char* ipdm_last_error(void) { return ipdm::last_error(); }
extern "C" int ipdm_abi_version(void) { return 1; }
extern "C" unsigned long long ipdm_launch_count(void) { return ipdm::g_launch_count; }
extern "C" void ipdm_launch_count_reset(void) { ipdm::g_launch_count = 0; }
