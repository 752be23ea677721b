// Engine<float>: production arithmetic (FMA contraction on).
#include "engine_dynamic.cuh"
namespace vx3b {
EngineBase *make_engine_f32() { return new EngineMux<float>(); }
} // namespace vx3b

#ifdef RES_PROFILE
// profile builds only (scripts/gpu_resident_phases.py): read (and clear) k_resident's per-phase warp-cycle counters
extern "C" int vx3b_debug_resident_profile(unsigned long long *out, int reset) {
    if (cudaMemcpyFromSymbol(out, vx3b::g_resProf, sizeof(unsigned long long) * 16) != cudaSuccess)
        return -1;
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(vx3b::g_resProf, z, sizeof(z));
    }
    return 0;
}
#endif
