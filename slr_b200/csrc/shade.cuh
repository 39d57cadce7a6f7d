// Surface points, light selection / sampling and emission for the shade stage.
//   Triangle::getSurfacePoint / sample / area         libSLR/Surface/TriangleMesh.cpp:180-259
//   BumpSingleSurfaceObject::getSurfacePoint          libSLR/Core/SurfaceObject.cpp:123-134
//   TransformedSurfaceObject (instances)              libSLR/Core/SurfaceObject.cpp:307-336,351-364, geometry.cpp:63-78
//   InfiniteSphere / InfiniteSphereSurfaceObject      libSLR/Surface/InfiniteSphere.cpp:34-59, SurfaceObject.cpp:158-185,217-222
//   Scene / aggregate light selection                 libSLR/Core/SurfaceObject.cpp:279-299,432-466
//   piecewise-constant distributions                  libSLR/Core/distributions.cpp:97-119,172-224
//   DiffuseEDF / IBLEDF                               libSLR/EDFs/basic_EDFs.cpp:12-29, IBLEDF.cpp:11-29
#pragma once
#include "material.cuh"
#include "motion.cuh"

namespace slrgpu {

struct TriVerts {
    float4 p0, n0, t0, p1, n1, t1, p2, n2, t2;    // SlrGpuVertex x 3: (pos,u) (normal,v) (tangent,-)
};

__device__ __forceinline__ TriVerts loadTriangle(const DeviceScene& s, const SlrGpuTriangle& tri) {
    TriVerts t;
    const float4* a = s.vertices + (size_t)tri.v[0] * 3;
    const float4* b = s.vertices + (size_t)tri.v[1] * 3;
    const float4* c = s.vertices + (size_t)tri.v[2] * 3;
    t.p0 = __ldg(a); t.n0 = __ldg(a + 1); t.t0 = __ldg(a + 2);
    t.p1 = __ldg(b); t.n1 = __ldg(b + 1); t.t1 = __ldg(b + 2);
    t.p2 = __ldg(c); t.n2 = __ldg(c + 1); t.t2 = __ldg(c + 2);
    return t;
}
__device__ __forceinline__ V3 xyz(const float4& v) { return V3(v.x, v.y, v.z); }

__device__ __forceinline__ float triangleArea(const TriVerts& t) {
    return 0.5f * length(cross(xyz(t.p1) - xyz(t.p0), xyz(t.p2) - xyz(t.p0)));
}

// shading frame at barycentrics (b0, b1, b2); `orthogonalize` is Triangle::getSurfacePoint's Gram-Schmidt
// step, which Triangle::sample does not have
__device__ __forceinline__ void triangleFrame(const TriVerts& t, float b0, float b1, float b2, bool orthogonalize, SurfPt* sp) {
    sp->gn = normalize(cross(xyz(t.p1) - xyz(t.p0), xyz(t.p2) - xyz(t.p0)));
    sp->tu = b0 * t.p0.w + b1 * t.p1.w + b2 * t.p2.w;
    sp->tv = b0 * t.n0.w + b1 * t.n1.w + b2 * t.n2.w;
    sp->sf.z = normalize(b0 * xyz(t.n0) + b1 * xyz(t.n1) + b2 * xyz(t.n2));
    sp->sf.x = normalize(b0 * xyz(t.t0) + b1 * xyz(t.t1) + b2 * xyz(t.t2));
    if (orthogonalize) {
        const float dotNT = dot(sp->sf.z, sp->sf.x);
        if (fabsf(dotNT) >= 0.01f) sp->sf.x = normalize(sp->sf.x - dotNT * sp->sf.z);
    }
    sp->sf.y = cross(sp->sf.z, sp->sf.x);
}

__device__ __forceinline__ void applyNormalMap(const DeviceScene& s, uint32_t normalMap, SurfPt* sp) {
    const V3 nLocal = evalNormalTexture(s, normalMap, *sp);
    const V3 tLocal = V3(1, 0, 0) - dot(nLocal, V3(1, 0, 0)) * nLocal;
    const V3 bLocal = V3(0, 1, 0) - dot(nLocal, V3(0, 1, 0)) * nLocal;
    const V3 t = normalize(sp->sf.fromLocal(tLocal));
    const V3 b = normalize(sp->sf.fromLocal(bLocal));
    const V3 n = normalize(sp->sf.fromLocal(nLocal));
    sp->sf.x = t; sp->sf.y = b; sp->sf.z = n;
}

// operator*(StaticTransform, SurfacePoint) (geometry.cpp:63-78): normal by the transposed inverse,
// the frame vectors as plain vectors, everything re-normalised
__device__ __forceinline__ void transformSurfPt(const InstanceXfm& x, SurfPt* sp) {
    sp->p = xfmPoint(x.mat, sp->p);
    sp->gn = normalize(xfmNormal(x.matInv, sp->gn));
    sp->sf.x = normalize(xfmVector(x.mat, sp->sf.x));
    sp->sf.y = normalize(xfmVector(x.mat, sp->sf.y));
    sp->sf.z = normalize(xfmVector(x.mat, sp->sf.z));
}

// Intersection -> SurfacePoint for a triangle hit. Returns the hit triangle's record; *localArea is
// the area evaluateAreaPDF uses (object space, also for instances -- as the reference). `time`: the ray's time, at which a
// moving instance's transform is sampled (TransformedSurfaceObject::getSurfacePoint, SurfaceObject.cpp:320-336).
static __device__ __noinline__ SlrGpuTriangle hitSurfacePoint(const DeviceScene& s, uint32_t prim, uint32_t inst, float t, float b0, float b1,
                                                 const V3& org, const V3& dir, float time, SurfPt* sp, float* localArea) {
    const SlrGpuTriangle tri = s.triangles[prim];
    const TriVerts tv = loadTriangle(s, tri);
    sp->atInfinity = false;
    sp->prim = prim; sp->inst = inst;
    sp->u = b0; sp->v = b1;
    triangleFrame(tv, b0, b1, 1.0f - b0 - b1, true, sp);
    *localArea = triangleArea(tv);
    if (inst == SLRGPU_INVALID_ID) {
        sp->p = org + dir * t;
        if (tri.normal_map != SLRGPU_INVALID_ID) applyNormalMap(s, tri.normal_map, sp);
    } else {
        float scratch[32];
        const InstanceXfm x = instanceTransformAt(s, s.instances[inst], time, scratch);
        const V3 lo = xfmPoint(x.matInv, org), ld = xfmVector(x.matInv, dir);
        sp->p = lo + ld * t;
        if (tri.normal_map != SLRGPU_INVALID_ID) applyNormalMap(s, tri.normal_map, sp);
        transformSurfPt(x, sp);
    }
    return tri;
}

// InfiniteSphere::intersect + getSurfacePoint for a ray that left the scene
__device__ inline void envSurfacePoint(const V3& dir, SurfPt* sp) {
    const float theta = acosf(fminf(fmaxf(dir.y, -1.0f), 1.0f));
    const float phi = fmodf(atan2f(-dir.x, dir.z) + 2 * kPi, 2 * kPi);
    sp->p = dir;
    sp->atInfinity = true;
    sp->gn = -dir;
    sp->u = phi; sp->v = theta;
    sp->tu = phi / (2 * kPi); sp->tv = theta / kPi;
    float sph, cph;
    sincosf(phi, &sph, &cph);
    sp->sf.x = V3(-cph, 0.0f, -sph);
    sp->sf.z = sp->gn;
    sp->sf.y = cross(sp->sf.z, sp->sf.x);
    sp->prim = SLRGPU_INVALID_ID; sp->inst = SLRGPU_INVALID_ID;
}

// ---- piecewise-constant distributions --------------------------------------------------------
__device__ __forceinline__ uint32_t prevPowerOf2(uint32_t x) {
    x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return x - (x >> 1);
}
// RegularConstantDiscrete1D::sample over an aggregate's light list (CDF[k] = cdf_lo of entry k)
__device__ inline uint32_t sampleLightList(const SlrGpuLight* __restrict__ lights, uint32_t n, float u, float* prob, float* remapped) {
    int idx = (int)n;
    for (int d = (int)prevPowerOf2(n); d > 0; d >>= 1)
        if (idx - d > 0 && lights[idx - d].cdf_lo >= u) idx -= d;
    --idx;
    const SlrGpuLight l = lights[idx];
    *prob = l.pmf;
    *remapped = (u - l.cdf_lo) / (l.cdf_hi - l.cdf_lo);
    return (uint32_t)idx;
}
// RegularConstantContinuous1D::sample over (pdf[n], cdf[n + 1])
__device__ inline float sampleContinuous1D(const float* __restrict__ pdf, const float* __restrict__ cdf, uint32_t n, float u, float* PDF) {
    int idx = (int)n;
    for (int d = (int)prevPowerOf2(n); d > 0; d >>= 1)
        if (idx - d > 0 && __ldg(cdf + idx - d) >= u) idx -= d;
    --idx;
    *PDF = __ldg(pdf + idx);
    const float c0 = __ldg(cdf + idx), c1 = __ldg(cdf + idx + 1);
    const float t = (u - c0) / (c1 - c0);
    return (idx + t) / n;
}
__device__ inline float envEvaluateUVPDF(const DeviceScene& s, float d0, float d1) {
    const uint32_t W = s.envMapWidth, H = s.envMapHeight;
    const uint32_t row = min((uint32_t)(H * d1), H - 1);
    const float top = __ldg(s.envMarginalPdf + min((uint32_t)(int32_t)(d1 * H), H - 1));
    return top * __ldg(s.envRowPdf + (size_t)row * W + min((uint32_t)(int32_t)(d0 * W), W - 1));
}

// ---- light sampling ----------------------------------------------------------------------------
struct LightSample {
    SurfPt sp;             // point on the light (world space)
    float lightPDF;        // selection probability x area pdf
    float areaPDF;
    uint32_t material;     // emitter material (for the emittance)
    bool isEnv;
};

// Scene::selectLight + Light::sample (SurfaceObject.cpp:432-452, 82-91, 158-185, 351-364)
static __device__ __noinline__ void sampleLight(const DeviceScene& s, float uSel, float u0, float u1, float time, LightSample* ls) {
    float prob = 1.0f;
    bool env = false;
    const float aggrImp = s.topLightImportance;
    if (s.envPresent) {
        const float sumImps = aggrImp + 1.0f;
        const float su = sumImps * uSel;
        if (su < aggrImp) { uSel = uSel / (aggrImp / sumImps); prob = aggrImp / sumImps; }
        else { env = true; prob = 1.0f / sumImps; }
    }
    ls->isEnv = env;
    if (env) {
        const uint32_t W = s.envMapWidth, H = s.envMapHeight;
        float topPDF, rowPDF;
        const float d1 = sampleContinuous1D(s.envMarginalPdf, s.envMarginalCdf, H, u1, &topPDF);
        const uint32_t row = min((uint32_t)(H * d1), H - 1);
        const float d0 = sampleContinuous1D(s.envRowPdf + (size_t)row * W, s.envRowCdf + (size_t)row * (W + 1), W, u0, &rowPDF);
        const float uvPDF = rowPDF * topPDF;
        const float phi = d0 * 2 * kPi, theta = d1 * kPi;
        float sph, cph, sth, cth;
        sincosf(phi, &sph, &cph);
        sincosf(theta, &sth, &cth);
        SurfPt& sp = ls->sp;
        sp.p = V3(-sph * sth, cth, cph * sth);
        sp.atInfinity = true;
        sp.gn = -sp.p;
        sp.u = phi; sp.v = theta;
        sp.tu = phi / (2 * kPi); sp.tv = theta / kPi;
        sp.sf.x = normalize(V3(-cph, 0.0f, -sph));
        sp.sf.z = sp.gn;
        sp.sf.y = cross(sp.sf.z, sp.sf.x);
        sp.prim = SLRGPU_INVALID_ID; sp.inst = SLRGPU_INVALID_ID;
        ls->areaPDF = uvPDF / (2 * kPi * kPi * sth);
        ls->lightPDF = prob * ls->areaPDF;
        ls->material = s.envMaterial;
        return;
    }
    float p1, rem;
    const uint32_t li = sampleLightList(s.lights, s.numTopLights, uSel, &p1, &rem);
    prob *= p1;
    uint32_t object = s.lights[li].object;
    uint32_t inst = SLRGPU_INVALID_ID;
    if (object & 0x80000000u) {
        inst = object & 0x7FFFFFFFu;
        const SlrGpuInstance& in = s.instances[inst];
        float p2, rem2;
        const uint32_t lj = sampleLightList(s.lights + in.light_base, in.num_lights, rem, &p2, &rem2);
        prob *= p2;
        object = s.lights[in.light_base + lj].object;
    }
    const SlrGpuTriangle tri = s.triangles[object];
    const TriVerts tv = loadTriangle(s, tri);
    const float su1 = sqrtf(u0);                          // uniformSampleTriangle (distributions.h:60-64)
    const float b0 = 1.0f - su1, b1 = u1 * su1, b2 = 1.0f - b0 - b1;
    SurfPt& sp = ls->sp;
    sp.p = b0 * xyz(tv.p0) + b1 * xyz(tv.p1) + b2 * xyz(tv.p2);
    sp.atInfinity = false;
    sp.u = b0; sp.v = b1;
    sp.prim = object; sp.inst = inst;
    triangleFrame(tv, b0, b1, b2, false, &sp);
    ls->areaPDF = 1.0f / triangleArea(tv);
    if (inst != SLRGPU_INVALID_ID) {
        // a light inside an instance: the instance's transform at the query's time (SurfaceObject.cpp:351-364)
        float scratch[32];
        transformSurfPt(instanceTransformAt(s, s.instances[inst], time, scratch), &sp);
    }
    ls->lightPDF = prob * ls->areaPDF;
    ls->material = tri.material;
}

// Scene::evaluateProb(Light(isect.obj)) for a hit on an emitter (SurfaceObject.cpp:454-466, 291-299, 345-350)
__device__ inline float lightSelectionProb(const DeviceScene& s, const SlrGpuTriangle& tri, uint32_t inst, bool isEnv) {
    const float aggrImp = s.topLightImportance;
    float scale = 1.0f;
    if (s.envPresent) {
        const float sumImps = aggrImp + 1.0f;
        if (isEnv) return 1.0f / sumImps;
        scale = aggrImp / sumImps;
    }
    if (inst == SLRGPU_INVALID_ID) return scale * s.lights[tri.light_index].pmf;
    const SlrGpuInstance& in = s.instances[inst];
    return scale * (s.lights[in.light_index].pmf * s.lights[in.light_base + tri.light_index].pmf);
}

}  // namespace slrgpu
