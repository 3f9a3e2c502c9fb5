// ref_cut.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE's own cutFace / cutCell classes.  oracle/build.py:build_ref_cut compiles this
// file together with the four reference files, read where they lie,
//     /root/reference/src/SimPLIC/cut/cutFace/cutFace.{H,C}
//     /root/reference/src/SimPLIC/cut/cutCell/cutCell.{H,C}
// against oracle/of_stub/OpenFOAMCutStub.H (a stand-in for the few OpenFOAM types they use; OpenFOAM itself is not
// installable here) into oracle/_ref/libref_cut.so.  Nothing of the reference is copied into this repository.
// The signatures mirror the geometry primitives of include/svof.h so tests drive all three libraries alike:
//   ref_cut_faces            <-> svof_cut_faces            (cutFace::calcSubFace,            cutFace.C:136-259)
//   ref_cut_cells            <-> svof_cut_cells            (cutCell::calcSubCell,            cutCell.C:343-542)
//   ref_find_signed_distance <-> svof_find_signed_distance (cutCell::findSignedDistance,     cutCell.C:611-799)
//   ref_face_fluxes          <-> svof_face_fluxes          (cutFace::timeIntegratedFaceFlux, cutFace.C:262-389)
//   ref_interface_points     <-> svof_plic_surface, per cell (cutCell::interfacePoints,      cutCell.C:545-608)
// Mesh geometry (Cf, C, V, magSf) and the face flatness are INPUTS here (they are OpenFOAM / reconstruction.C
// quantities, outside the four files); tests pass the arrays svof_get_field returns.
#include <cstdint>
#include <new>

#include "cutCell.H"

#include "../include/svof.h"

const Foam::vector Foam::vector::zero;

namespace
{
using namespace Foam;
typedef geometricVofExt::SimPLIC::cutFace RefCutFace;
typedef geometricVofExt::SimPLIC::cutCell RefCutCell;

struct RefCut
{
    fvMesh mesh;
    scalarField flat;
    scalarField alpha;
    vectorField iN, iC, iS;
    scalarField iD;
    volScalarField alphaF, iDF;
    volVectorField iNF, iCF, iSF;
    RefCutFace* cf;
    RefCutCell* cc;
    RefCut() : cf(nullptr), cc(nullptr) {}
    ~RefCut()
    {
        delete cc;
        delete cf;
    }
};

inline vector v3(const double* p, int64_t i) { return vector(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline void put3(double* p, int64_t i, const vector& v)
{
    p[3 * i] = v.x();
    p[3 * i + 1] = v.y();
    p[3 * i + 2] = v.z();
}
}  // namespace

extern "C" {

void* ref_cut_create(const svof_mesh* m, const double* Cf, const double* C, const double* V, const double* magSf,
                     const double* flatness)
{
    if (!m || !Cf || !C || !V || !magSf || !flatness) return nullptr;
    RefCut* r = new (std::nothrow) RefCut;
    if (!r) return nullptr;
    fvMesh& fm = r->mesh;
    fm.points_.setSize(m->n_points);
    for (int i = 0; i < m->n_points; ++i) fm.points_[i] = v3(m->points, i);
    fm.faces_.setSize(m->n_faces);
    fm.owner_.setSize(m->n_faces);
    fm.faceCentres_.setSize(m->n_faces);
    fm.magSf_.setSize(m->n_faces);
    r->flat.setSize(m->n_faces);
    for (int f = 0; f < m->n_faces; ++f)
    {
        face& fa = fm.faces_[f];
        for (int k = m->face_offsets[f]; k < m->face_offsets[f + 1]; ++k) fa.append(m->face_points[k]);
        fm.owner_[f] = m->owner[f];
        fm.faceCentres_[f] = v3(Cf, f);
        fm.magSf_[f] = magSf[f];
        r->flat[f] = flatness[f];
    }
    // primitiveMesh::calcCells: every face goes to its owner in ascending face order, then every internal face to
    // its neighbour in ascending face order
    fm.cells_.setSize(m->n_cells);
    for (int f = 0; f < m->n_faces; ++f) fm.cells_[m->owner[f]].append(f);
    for (int f = 0; f < m->n_internal_faces; ++f) fm.cells_[m->neighbour[f]].append(f);
    fm.cellCentres_.setSize(m->n_cells);
    fm.cellVolumes_.setSize(m->n_cells);
    fm.cellPoints_.setSize(m->n_cells);
    for (int c = 0; c < m->n_cells; ++c)
    {
        fm.cellCentres_[c] = v3(C, c);
        fm.cellVolumes_[c] = V[c];
        fm.cellPoints_[c] = fm.cells_[c].labels(fm.faces_);
    }
    r->alpha.setSize(m->n_cells);
    r->iN.setSize(m->n_cells);
    r->iC.setSize(m->n_cells);
    r->iS.setSize(m->n_cells);
    r->iD.setSize(m->n_cells);
    r->alphaF = volScalarField(r->alpha);
    r->iDF = volScalarField(r->iD);
    r->iNF = volVectorField(r->iN);
    r->iCF = volVectorField(r->iC);
    r->iSF = volVectorField(r->iS);
    r->cf = new RefCutFace(fm, r->flat);
    r->cc = new RefCutCell(fm, r->flat, r->alphaF, r->iNF, r->iDF, r->iCF, r->iSF);
    return r;
}

void ref_cut_destroy(void* h) { delete static_cast<RefCut*>(h); }

int ref_cut_faces(void* h, int32_t n_polys, int32_t n_verts, const double* pts, const double* normals, const double* dists,
                  int32_t* status, double* centres, double* areas)
{
    RefCut* r = static_cast<RefCut*>(h);
    pointField fPts(n_verts);
    for (int64_t i = 0; i < n_polys; ++i)
    {
        for (int k = 0; k < n_verts; ++k) fPts[k] = v3(pts, i * n_verts + k);
        status[i] = r->cf->calcSubFace(fPts, v3(normals, i), dists[i]);
        put3(centres, i, r->cf->subFaceCentre());
        put3(areas, i, r->cf->subFaceArea());
    }
    return 0;
}

// splitWarpedFace = false only: with true, calcSubCell reads the local polyhedron that only findSignedDistance builds
int ref_cut_cells(void* h, int32_t n, const int32_t* cells, const double* normals, const double* dists, int32_t* status,
                  double* vof, double* sub_volume, double* iface_centre, double* iface_area)
{
    RefCut* r = static_cast<RefCut*>(h);
    for (int64_t i = 0; i < n; ++i)
    {
        status[i] = r->cc->calcSubCell(cells[i], v3(normals, i), dists[i], false);
        vof[i] = r->cc->volumeOfFluid();
        sub_volume[i] = r->cc->subCellVolume();
        put3(iface_centre, i, r->cc->interfaceCentre());
        put3(iface_area, i, r->cc->interfaceArea());
    }
    return 0;
}

int ref_find_signed_distance(void* h, int32_t n, const int32_t* cells, const double* alphas, const double* normals,
                             int32_t split_warped_face, int32_t* status, double* dists, double* iface_centre, double* iface_area)
{
    RefCut* r = static_cast<RefCut*>(h);
    for (int64_t i = 0; i < n; ++i)
    {
        const int c = cells[i];
        r->alpha[c] = alphas[i];
        r->iN[c] = v3(normals, i);
        r->iD[c] = 0;
        r->iC[c] = vector();
        r->iS[c] = vector();
        status[i] = r->cc->findSignedDistance(c, split_warped_face != 0);
        dists[i] = r->iD[c];
        put3(iface_centre, i, r->iC[c]);
        put3(iface_area, i, r->iS[c]);
    }
    return 0;
}

int ref_face_fluxes(void* h, int32_t n, const int32_t* faces, const double* normals, const double* dists, const double* Un0,
                    double dt, const double* phi, double* dVf)
{
    RefCut* r = static_cast<RefCut*>(h);
    for (int64_t i = 0; i < n; ++i)
    {
        const int f = faces[i];
        dVf[i] = r->cf->timeIntegratedFaceFlux(f, v3(normals, i), dists[i], Un0[i], dt, phi[i], r->mesh.magSf_[f]);
    }
    return 0;
}

// cutCell::interfacePoints() of one cell cut by a given plane; returns the number of points (<= cap written)
int ref_interface_points(void* h, int32_t cell, const double* normal, double dist, int32_t cap, double* pts)
{
    RefCut* r = static_cast<RefCut*>(h);
    r->cc->calcSubCell(cell, v3(normal, 0), dist, false);
    const DynamicList<point>& ip = r->cc->interfacePoints();
    for (int k = 0; k < ip.size() && k < cap; ++k) put3(pts, k, ip[k]);
    return ip.size();
}

}  // extern "C"
