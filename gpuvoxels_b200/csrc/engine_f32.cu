// Engine<float>: production arithmetic (FMA contraction on).
#include "engine_impl.cuh"
namespace vx3b {
EngineBase *make_engine_f32() { return new Engine<float>(); }
} // namespace vx3b
