// Extracts the published physical data tables the spectral path needs from the compiled reference
// into one binary blob (slr_b200/data/spectral_tables.bin): CIE 1931 2-degree colour-matching
// functions, CIE D65, the 24 ColorChecker reflectances, the refractiveindex.info IOR tables and the
// Meng-Simon-Hanika-Dachsbacher 2015 spectral up-sampling grid. These are measurement data, not
// program logic; the blob is produced from the reference build so the product's tables are
// bit-identical to what libSLR computes with (BasicTypes/Spectrum.cpp:34-218, Spectrum.h:205-571,
// common_spectra.cpp, spectrum_library.cpp). Run: oracle/_ref/dump_tables <out.bin>
#include <libSLR/BasicTypes/Spectrum.h>
#include <libSLR/BasicTypes/SpectrumTypes.h>
#include "table_io.h"

int main(int argc, char** argv) {
    using namespace SLR;
    if (argc < 2) { fprintf(stderr, "usage: dump_tables out.bin\n"); return 2; }
    TableWriter w(argv[1]);
    w.floats("cmf/xbar_2deg", xbar_2deg, NumCMFSamples);
    w.floats("cmf/ybar_2deg", ybar_2deg, NumCMFSamples);
    w.floats("cmf/zbar_2deg", zbar_2deg, NumCMFSamples);
    w.bytes("upsampling/grid", Upsampling::spectrum_grid, sizeof(Upsampling::spectrum_grid));
    w.floats("upsampling/points", (const float*)Upsampling::spectrum_data_points, sizeof(Upsampling::spectrum_data_points) / 4);
    w.floats("illuminant/D65", StandardIlluminant::D65, StandardIlluminant::NumSamples);
    w.floats("colorchecker/spectra", &ColorChecker::Spectra[0][0], 24 * ColorChecker::NumSamples);
    for (const auto& kv : SpectrumLibrary::IORs) {
        const SpectrumLibrary::IndexOfRefraction& ior = kv.second;
        float meta[4] = {(float)(uint32_t)ior.dType, (float)ior.numSamples, ior.minLambdas, ior.maxLambdas};
        w.floats("ior/" + kv.first + "/meta", meta, 4);
        if (ior.lambdas) w.floats("ior/" + kv.first + "/lambdas", ior.lambdas, ior.numSamples);
        if (ior.etas) w.floats("ior/" + kv.first + "/etas", ior.etas, ior.numSamples);
        if (ior.ks) w.floats("ior/" + kv.first + "/ks", ior.ks, ior.numSamples);
    }
    // derived tables, for cross-checking the host's own initSpectrum restatement
    initSpectrum();
    w.floats("derived/xbar16", DiscretizedSpectrum::xbar.get(), 16);
    w.floats("derived/ybar16", DiscretizedSpectrum::ybar.get(), 16);
    w.floats("derived/zbar16", DiscretizedSpectrum::zbar.get(), 16);
    w.floats("derived/integralCMF", &DiscretizedSpectrum::integralCMF, 1);
    return 0;
}
