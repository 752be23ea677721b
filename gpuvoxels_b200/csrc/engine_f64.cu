// Engine<double>: validation build. Compiled with -fmad=false so every double operation rounds exactly as
// the reference CPU library's does; only libm calls (sin, cos, acos) can differ, by an ulp.
#include "engine_dynamic.cuh"
namespace vx3b {
EngineBase *make_engine_f64() { return new EngineMux<double>(); }
} // namespace vx3b
