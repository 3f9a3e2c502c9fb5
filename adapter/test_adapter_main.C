// Driver for tests/test_adapter.py: builds a (stub) fvMesh and fields from a binary dump written by the test, runs the
// adapter class the way plicVofAdvectionFoam does (plicVof.H:13-57), and dumps alpha1, alphaPhi and rhoPhi.
//   test_adapter_main <in.bin> <out.bin> <nSteps>
#include <cstdio>
#include <cstdlib>

#include "solveVofEquB200.H"

using namespace Foam;

template<class T>
static void rd(FILE* f, T* p, size_t n)
{
    if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv)
{
    if (argc < 4) return 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    int32_t hdr[6];
    rd(f, hdr, 6);
    const label nP = hdr[0], nF = hdr[1], nIF = hdr[2], nC = hdr[3], nPatches = hdr[4], nFP = hdr[5];
    fvMesh mesh;
    mesh.nCells_ = nC;
    mesh.points_.setSize(nP);
    rd(f, reinterpret_cast<double*>(mesh.points_.data()), size_t(3)*nP);
    labelList off(nF + 1), pts(nFP);
    rd(f, off.data(), size_t(nF) + 1);
    rd(f, pts.data(), size_t(nFP));
    mesh.faces_.setSize(nF);
    forAll(mesh.faces_, i)
    {
        mesh.faces_[i].setSize(off[i + 1] - off[i]);
        forAll(mesh.faces_[i], k) mesh.faces_[i][k] = pts[off[i] + k];
    }
    mesh.owner_.setSize(nF);
    mesh.neighbour_.setSize(nIF);
    rd(f, mesh.owner_.data(), size_t(nF));
    rd(f, mesh.neighbour_.data(), size_t(nIF));
    mesh.patches_.setSize(nPatches);
    forAll(mesh.patches_, pi)
    {
        int32_t ps[3];
        rd(f, ps, 3);
        mesh.patches_[pi].start_ = ps[0];
        mesh.patches_[pi].size_ = ps[1];
        mesh.patches_[pi].type_ = ps[2] == 1 ? "empty" : "wall";
        mesh.patches_[pi].name_ = "patch" + std::to_string(pi);
    }
    mesh.Cf_.setSize(nF); mesh.Sf_.setSize(nF); mesh.C_.setSize(nC); mesh.V_.setSize(nC);
    rd(f, reinterpret_cast<double*>(mesh.Cf_.data()), size_t(3)*nF);
    rd(f, reinterpret_cast<double*>(mesh.Sf_.data()), size_t(3)*nF);
    rd(f, reinterpret_cast<double*>(mesh.C_.data()), size_t(3)*nC);
    rd(f, mesh.V_.data(), size_t(nC));
    double dt;
    rd(f, &dt, 1);
    mesh.time_.deltaT_ = dt;
    // fvSolution solvers."alpha.water" of tutorials/test/plicVofAdvectionFoam/system/fvSolution:21-35
    dictionary& d = mesh.solverDicts_["alpha.water"];
    d.add("nAlphaBounds", "3"); d.add("snapTol", "0"); d.add("clip", "false"); d.add("mixedCellTol", "1e-8");
    d.add("orientationMethod", "LS"); d.add("splitWarpedFace", "false"); d.add("writePlicFields", "true");
    d.add("nAlphaSubCycles", "1"); d.add("cAlpha", "1"); d.add("tolerance", "1e-9");   // the last one is not ours: ignored

    volScalarField alpha1("alpha.water", mesh);
    forAll(mesh.patches_, pi) alpha1.boundaryFieldRef()[pi].type_ = "zeroGradient";
    rd(f, alpha1.primitiveFieldRef().data(), size_t(nC));
    alpha1.correctBoundaryConditions();
    surfaceScalarField phi("phi", mesh);
    scalarField phiFlat(nF);
    rd(f, phiFlat.data(), size_t(nF));
    forAll(phi.primitiveFieldRef(), i) phi.primitiveFieldRef()[i] = phiFlat[i];
    forAll(mesh.patches_, pi)
    {
        forAll(phi.boundaryFieldRef()[pi], i) phi.boundaryFieldRef()[pi][i] = phiFlat[mesh.patches_[pi].start_ + i];
    }
    volVectorField U("U", mesh);
    rd(f, reinterpret_cast<double*>(U.primitiveFieldRef().data()), size_t(3)*nC);
    fclose(f);

    geometricVofExt::SimPLIC::solveVofEqu plicVofSolver(alpha1, phi, U);     // createFields.H:138
    const int nSteps = atoi(argv[3]);
    for (int k = 0; k < nSteps; ++k)
    {
        plicVofSolver.reconstruct();                                          // plicVof.H:37
        plicVofSolver.advect(zeroField(), zeroField());                       // plicVof.H:41
    }
    tmp<surfaceScalarField> rhoPhi(plicVofSolver.getRhoPhi(dimensionedScalar("rho1", 1000.0), dimensionedScalar("rho2", 1.0)));
    volScalarField rho1("rho1", mesh, 1000.0), rho2("rho2", mesh, 1.0);
    forAll(mesh.patches_, pi)
    {
        rho1.boundaryFieldRef()[pi] = scalarField(mesh.patches_[pi].size_, 1000.0);
        rho2.boundaryFieldRef()[pi] = scalarField(mesh.patches_[pi].size_, 1.0);
    }
    tmp<surfaceScalarField> rhoPhi2(plicVofSolver.getRhoPhi(rho1, rho2));
    // the registry contract of the samplers (sampledPlicSurface.C:96-99), on the interface of the final field
    plicVofSolver.reconstruct();
    geometricVofExt::SimPLIC::plicSurface surf =
        mesh.lookupObjectRef<geometricVofExt::SimPLIC::reconstruction>("reconstruction").interface();
    Info<< "plicSurface: " << surf.faces.size() << " polygons, " << surf.points.size() << " points; reconstructionTime "
        << plicVofSolver.reconstructionTime() << " s, advectionTime " << plicVofSolver.advectionTime() << " s" << endl;

    // the second sampler's source (sampledReconstructedSubcellFaces.C:97-100)
    geometricVofExt::SimPLIC::plicSurface sub =
        mesh.lookupObjectRef<geometricVofExt::SimPLIC::reconstruction>("reconstruction").subCellFaces();
    Info<< "subCellFaces: " << sub.faces.size() << " faces, " << sub.points.size() << " points" << endl;

    FILE* o = fopen(argv[2], "wb");
    fwrite(alpha1.primitiveField().cdata(), sizeof(double), size_t(nC), o);
    fwrite(plicVofSolver.alphaPhi().primitiveField().cdata(), sizeof(double), size_t(nIF), o);
    fwrite(rhoPhi().primitiveField().cdata(), sizeof(double), size_t(nIF), o);
    fwrite(rhoPhi2().primitiveField().cdata(), sizeof(double), size_t(nIF), o);
    const int32_t nPoly = int32_t(surf.faces.size());
    fwrite(&nPoly, sizeof(int32_t), 1, o);
    const int32_t nSub = int32_t(sub.faces.size());
    fwrite(&nSub, sizeof(int32_t), 1, o);
    fclose(o);

    // overset hook (reconstruction.C:649-662): with cell types set, only CALCULATED cells are listed
    geometricVofExt::SimPLIC::reconstruction& rec = mesh.lookupObjectRef<geometricVofExt::SimPLIC::reconstruction>("reconstruction");
    const label nAll = label(rec.mixedCells().size());
    labelList types(nC, 0);
    for (label c = 0; c < nC; c += 2) types[c] = 1;      // every other cell INTERPOLATED
    plicVofSolver.setCellTypes(types);
    plicVofSolver.reconstruct();
    const labelList kept(rec.mixedCells());
    bool onlyCalculated = true;
    forAll(kept, i) onlyCalculated = onlyCalculated && (types[kept[i]] == 0);
    plicVofSolver.setCellTypes(labelList());
    plicVofSolver.reconstruct();
    Info<< "overset filter: " << label(kept.size()) << " of " << nAll << " interface cells kept, CALCULATED only "
        << (onlyCalculated ? "yes" : "no") << ", restored " << label(rec.mixedCells().size()) << endl;
    return 0;
}
