// AnimatedTransform through the REFERENCE (test infrastructure only): decomposes two key matrices and samples the
// transform at the given times with the reference's own classes (libSLR/Core/Transform.h:89-144,
// BasicTypes/Quaternion.cpp:15-43), and its motionBounds of a box.
//   ref_motion in.bin out.bin
// in.bin : f32 matBegin[16], matEnd[16] (column-major), tBegin, tEnd, box lo[3], hi[3], u32 n, f32 times[n]
// out.bin: f32 T0[3] R0[4] S0[16] T1[3] R1[4] S1[16], motionBounds lo[3] hi[3], then n x (mat[16], matInv[16])
#include <libSLR/Core/Transform.h>
#include <libSLR/BasicTypes/Quaternion.h>
#include <cstdio>
#include <vector>

using namespace SLR;

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_motion in.bin out.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    float mb[16], me[16], tb, te, box[6];
    uint32_t n = 0;
    if (fread(mb, 4, 16, f) != 16 || fread(me, 4, 16, f) != 16 || fread(&tb, 4, 1, f) != 1 || fread(&te, 4, 1, f) != 1 ||
        fread(box, 4, 6, f) != 6 || fread(&n, 4, 1, f) != 1) return 1;
    std::vector<float> times(n);
    if (n && fread(times.data(), 4, n, f) != n) return 1;
    fclose(f);
    StaticTransform b{Matrix4x4(mb)}, e{Matrix4x4(me)};
    AnimatedTransform anim(b, e, tb, te);
    f = fopen(argv[2], "wb");
    for (int k = 0; k < 2; ++k) {
        fwrite(&anim.m_T[k], 4, 3, f);
        const float q[4] = {anim.m_R[k].x, anim.m_R[k].y, anim.m_R[k].z, anim.m_R[k].w};
        fwrite(q, 4, 4, f);
        fwrite(&anim.m_S[k], 4, 16, f);
    }
    const BoundingBox3D mbounds = anim.motionBounds(BoundingBox3D(Point3D(box[0], box[1], box[2]), Point3D(box[3], box[4], box[5])));
    fwrite(&mbounds.minP, 4, 3, f); fwrite(&mbounds.maxP, 4, 3, f);
    for (uint32_t i = 0; i < n; ++i) {
        StaticTransform tf;
        anim.sample(times[i], &tf);
        fwrite(&tf.mat, 4, 16, f);
        fwrite(&tf.matInv, 4, 16, f);
    }
    fclose(f);
    return 0;
}
