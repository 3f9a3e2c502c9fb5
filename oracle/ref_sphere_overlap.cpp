// TEST INFRASTRUCTURE ONLY.
//
// Thin C wrapper (ours) around the reference's vendored exact sphere/hexahedron
// overlap library, compiled FROM WHERE IT LIES under /root/reference:
//   applications/test/calcExactVofFieldForSphericalShapeInHexMesh/overlap.hpp (+ Eigen/)
// It mirrors functions.H:1-27 (cutVolume) of that app.  The compiled object goes
// to oracle/_ref/ (git-ignored, travels to the GPU box); no reference source is
// copied into this repository.  Used by tests to build the exact initial
// volume-fraction field of the LeVeque sphere.
#include "overlap.hpp"

extern "C" int ref_sphere_hex_overlap(long n, const double* hexes /* [n][8][3] */, const double* centre, double radius,
                                      double* volume /* [n] */)
{
    const Sphere shape{vector_t(centre[0], centre[1], centre[2]), radius};
    for (long i = 0; i < n; ++i) {
        const double* q = hexes + i * 24;
        vector_t v[8];
        for (int k = 0; k < 8; ++k) v[k] = vector_t(q[3 * k], q[3 * k + 1], q[3 * k + 2]);
        Hexahedron hexI{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
        volume[i] = double(overlap(shape, hexI));
    }
    return 0;
}
