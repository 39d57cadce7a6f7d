// Renderer entry points (placeholder until the wavefront path tracer lands).
#include "device_scene.h"
using namespace slrgpu;
extern "C" {
SLRGPU_API int slrgpu_render(SlrGpuScene*, const SlrGpuRenderParams*, float*, SlrGpuRenderStats*) {
    setError("slrgpu_render: not implemented in this build");
    return SLRGPU_ERR_UNSUPPORTED;
}
SLRGPU_API int slrgpu_render_device(SlrGpuScene*, const SlrGpuRenderParams*, float*, void*, SlrGpuRenderStats*) {
    setError("slrgpu_render_device: not implemented in this build");
    return SLRGPU_ERR_UNSUPPORTED;
}
}
