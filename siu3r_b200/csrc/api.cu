// Process-wide bookkeeping shared by all C-ABI entry points.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

static std::atomic<long long> g_launches{0};

extern "C" {

void siu3r_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Number of kernels of this library launched since the last reset (bench.py -> "gpu_launches").
long long siu3r_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void siu3r_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

// Programmatic dependent launch switch (common.cuh): SIU3R_PDL=0 in the environment, or siu3r_set_pdl(0), turns the launch attribute off.
static int g_pdl = -1;
int siu3r_pdl_enabled(void) {
    if (g_pdl < 0) { const char* e = getenv("SIU3R_PDL"); g_pdl = (e && e[0] == '0') ? 0 : 1; }
    return g_pdl;
}
void siu3r_set_pdl(int on) { g_pdl = on ? 1 : 0; }

// ABI version of include/siu3r_b200.h this library was built against.
int siu3r_abi_version(void) { return 1; }

}  // extern "C"
