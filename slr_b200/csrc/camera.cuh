// The camera sample of a path: shared by the path tracer's ray generation (render.cu), the debug renderer and the eye
// subpath of the bidirectional path tracer (bpt.cu).
//   PerspectiveCamera::sample / PerspectiveIDF::sample   libSLR/Cameras/PerspectiveCamera.cpp:33-74
//   Camera::sampleRay                                    libSLR/Core/cameras.h:46-57
//   WavelengthSamples::createWithEqualOffsets            libSLR/BasicTypes/SpectrumTypes.h:54-64
#pragma once
#include "rng.cuh"
#include "shade.cuh"
#include "wavefront.cuh"

namespace slrgpu {

// One camera sample of pixel (x, y): Job::kernel's first half (PathTracingRenderer.cpp:100-126; DebugRenderer.cpp:134-152
// draws the same samples) -- time, jittered pixel position, wavelengths, lens position, PerspectiveCamera::sample and
// PerspectiveIDF::sample.
struct CameraSample {
    V3 org, dir;
    float weight, wlOffset, time;
    uint32_t ipx, ipy, hero, flags;
    // what the bidirectional path tracer's first eye vertex keeps (LensPosQueryResult::surfPt, IDFQueryResult)
    Frame lensFrame;            // z = lens normal
    float lensU, lensV;         // concentric disk sample (surfPt.u, surfPt.v)
    float dirLocalZ, dirPDF;
    float px, py;               // float pixel position of the sample
};
template <int NC>
__device__ __forceinline__ void sampleCamera(const DeviceScene& s, const RenderConstants& rc, uint32_t x, uint32_t y, uint32_t pixel, uint32_t sample,
                                             CameraSample* o) {
    const Rand4 r0 = pathRandom(rc.seed, pixel, sample, 0);    // time, pixel x, pixel y, wavelength offset
    const Rand4 r1 = pathRandom(rc.seed, pixel, sample, 1);    // wavelength selection, lens u0, lens u1
    const float px = x + r0.y, py = y + r0.z;
    const float wlOffset = r0.w;
    o->wlOffset = wlOffset;
    // IndependentLightPathSampler::getTimeSample (light_path_samplers.h:50)
    const float time = rc.timeStart * (1 - r0.x) + rc.timeEnd * r0.x;
    o->time = time;
    o->hero = min((uint32_t)(NC * r1.x), (uint32_t)(NC - 1));

    // PerspectiveCamera::sample
    float lx, ly;
    concentricSampleDisk(r1.y, r1.z, &lx, &ly);
    const SlrGpuCamera& cam = s.camera;
    // the camera's transform at the sample's time (PerspectiveCamera::sample, PerspectiveCamera.cpp:34-36)
    float camScratch[32];
    const float* camMat = cam.mat;
    const float* camInv = cam.mat_inv;
    if (s.cameraMotion != 0u && s.motions != nullptr) {
        sampleMotion(s.motions[s.cameraMotion - 1u], cam.mat, cam.mat_inv, time, camScratch, camScratch + 16);
        camMat = camScratch; camInv = camScratch + 16;
    }
    const V3 orgLocal(cam.lens_radius * lx, cam.lens_radius * ly, 0.0f);
    o->org = xfmPoint(camMat, orgLocal);
    const V3 lensN = xfmNormal(camInv, V3(0, 0, 1));
    Frame f;
    f.z = lensN;
    f.x = xfmVector(camMat, V3(1, 0, 0));
    f.y = cross(f.z, f.x);
    // PerspectiveIDF::sample with (p.x / W, p.y / H)
    const V3 pFocus(rc.opWidth * (0.5f - px / rc.width), rc.opHeight * (0.5f - py / rc.height), cam.obj_plane_dist);
    const V3 dirLocal = normalize(pFocus - orgLocal);
    const float dirPDF = cam.img_plane_dist * cam.img_plane_dist / ((dirLocal.z * dirLocal.z * dirLocal.z) * rc.imgPlaneArea);
    const V3 dir = f.fromLocal(dirLocal);
    o->dir = dir;
    o->lensFrame = f; o->lensU = lx; o->lensV = ly; o->dirLocalZ = dirLocal.z; o->dirPDF = dirPDF; o->px = px; o->py = py;
    o->weight = absDot(dir, lensN) / (rc.lensAreaPDF * dirPDF * rc.selectWLPDF);

    // ImageSensor::add bins by the float pixel position
    o->ipx = min((uint32_t)px, rc.width - 1); o->ipy = min((uint32_t)py, rc.height - 1);
    uint32_t flags = kFlagCameraRay;
    // wavelength i lands in stratum i unless the offset sits within rounding distance of 0 or 1:
    // only then is the exact (16 x 2 IEEE divisions) test needed
    if (NC == 16 && ((wlOffset > 1e-4f && wlOffset < 1.0f - 1e-4f) || strataInPlace(wlOffset))) flags |= kFlagStrataInPlace;
    o->flags = flags;
}

}  // namespace slrgpu
