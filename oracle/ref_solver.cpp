// ref_solver.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE's own solveVofEqu class.  oracle/build.py:build_ref_solver compiles this file
// together with the reference files, read where they lie,
//     /root/reference/src/SimPLIC/solveVofEqu/solveVofEqu.{H,C}, solveVofEquTemplates.C
//     /root/reference/src/SimPLIC/reconstruction/reconstruction.{H,C}
//     /root/reference/src/SimPLIC/advection/advection.{H,C}, advectionTemplates.C
//     /root/reference/src/SimPLIC/cut/cutFace/cutFace.{H,C}, cut/cutCell/cutCell.{H,C}
// (every file of src/SimPLIC outside sampling/) against oracle/of_stub_rec/ (a stand-in for the OpenFOAM types they
// use; OpenFOAM itself is not installable here) into oracle/_ref/libref_solver.so.  Nothing of the reference is copied
// into this repository.
//
//   ref_solver_create       <->  svof_create          (solveVofEqu ctor: reconstruction ctor incl. updateFaceFlatness,
//                                                      advection ctor)
//   ref_solver_set_state    <->  svof_set_alpha/phi/U
//   ref_solver_reconstruct  <->  svof_reconstruct     (solveVofEqu::reconstruct, reconstruction.C:680-722)
//   ref_solver_advect       <->  svof_advect          (solveVofEqu::advect, advectionTemplates.C:352-418)
//   ref_solver_map_alpha    <->  svof_map_alpha_field (reconstruction::mapAlphaField, reconstruction.C:725-784)
//   ref_solver_surface      <->  svof_plic_surface / svof_subcell_faces (reconstruction::interface / subCellFaces)
//   ref_solver_set_cell_types    the overset filter of reconstruction::initialize (reconstruction.C:649-662)
//
// Inputs that are OpenFOAM's, not the reference's: mesh geometry (from the oracle's mesh services), the patch-field
// evaluation of alpha (written here as the oracle writes it), interpolationCellPoint and the cell-point-cell stencil
// membership (answered by the oracle's restatements), the LU solve of the 4 x 4 least-squares system.
#include <cstdint>
#include <map>
#include <memory>
#include <new>
#include <string>

#include "ora_solver.hpp"   // before the stand-in: it defines the macros Info / forAll

#define NoRepository
#include "solveVofEqu.H"

const Foam::vector Foam::vector::zero;

namespace Foam
{
std::ostringstream& stubInfoStream()
{
    static std::ostringstream s;
    return s;
}
StubFatalError FatalError;
std::function<vector(const vector&, label)> stubInterpolateCellPoint;
std::function<void(const fvMesh&, label, labelList&)> stubCpcStencil;
int stubGradScheme = 0;
dictionary& stubDynamicMeshDict()
{
    static dictionary d;
    return d;
}
void stubLUsolve4(scalar A[4][4], scalar b[4], int m) { ora::LUsolve(A, b, m); }

static std::map<const fvMesh*, std::unique_ptr<zoneDistribute>> zoneDistributes;
zoneDistribute& zoneDistribute::New(const fvMesh& mesh)
{
    std::unique_ptr<zoneDistribute>& p = zoneDistributes[&mesh];
    if (!p) p.reset(new zoneDistribute(mesh));
    return *p;
}
static std::map<const fvMesh*, cellCellStencilObject> oversetStencils;
cellCellStencilObject& stubOversetStencil(const fvMesh& mesh) { return oversetStencils[&mesh]; }
}  // namespace Foam

namespace
{
using namespace Foam;
typedef geometricVofExt::SimPLIC::solveVofEqu RefSolver;

struct Ref
{
    ora::Solver S;   // mesh services (geometry, stencil membership, U interpolation); its own solver is never called
    std::unique_ptr<fvMesh> fmp;
    volScalarField* alpha;
    surfaceScalarField* phi;
    volVectorField* U;
    volScalarField* cellMask;   // overset: 0 in HOLE cells, 1 elsewhere (registered as "cellMask")
    RefSolver* solver;
    DimensionedField<scalar> Sp, Su;
    std::string log, ctorLog, err;
    // last extracted surface
    std::vector<double> sPts;
    std::vector<int32_t> sOff, sFacePts, sCells;
    Ref() : alpha(nullptr), phi(nullptr), U(nullptr), cellMask(nullptr), solver(nullptr) {}
    ~Ref()
    {
        delete solver;
        if (fmp)
        {
            zoneDistributes.erase(fmp.get());
            oversetStencils.erase(fmp.get());
        }
        delete cellMask;
        delete U;
        delete phi;
        delete alpha;
    }
};

inline vector v3(const double* p, int64_t i) { return vector(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline vector v3(const ora::vec& v) { return vector(v.x, v.y, v.z); }

void setSurface(surfaceScalarField& f, const fvMesh& fm, const double* a)
{
    for (label i = 0; i < fm.nInternalFaces(); ++i) f[i] = a[i];
    for (label p = 0; p < f.boundaryField().size(); ++p)
    {
        const label start = fm.boundaryMesh()[p].start();
        for (label i = 0; i < f.boundaryField()[p].size(); ++i) f.boundaryFieldRef()[p][i] = a[start + i];
    }
}
void getSurface(const surfaceScalarField& f, const fvMesh& fm, double* a)
{
    for (label i = 0; i < fm.nFaces(); ++i) a[i] = 0.0;   // empty patches carry no field
    for (label i = 0; i < fm.nInternalFaces(); ++i) a[i] = f[i];
    for (label p = 0; p < f.boundaryField().size(); ++p)
    {
        const label start = fm.boundaryMesh()[p].start();
        for (label i = 0; i < f.boundaryField()[p].size(); ++i) a[start + i] = f.boundaryField()[p][i];
    }
}
template<class SurfaceVectorField>
void setSurfaceVec(SurfaceVectorField& f, const fvMesh& fm, const std::vector<ora::vec>& a)
{
    for (label i = 0; i < fm.nInternalFaces(); ++i) f[i] = v3(a[i]);
    for (label p = 0; p < f.boundaryField().size(); ++p)
    {
        const label start = fm.boundaryMesh()[p].start();
        for (label i = 0; i < f.boundaryField()[p].size(); ++i) f.boundaryFieldRef()[p][i] = v3(a[start + i]);
    }
}

void fillMesh(fvMesh& fm, const svof_mesh* m, const ora::Mesh& om, const int geomD[3])
{
    fm.nInternalFaces_ = m->n_internal_faces;
    fm.points_.setSize(m->n_points);
    for (int i = 0; i < m->n_points; ++i) fm.points_[i] = v3(m->points, i);
    fm.faces_.setSize(m->n_faces);
    fm.owner_.setSize(m->n_faces);
    fm.neighbour_.setSize(m->n_internal_faces);
    fm.faceCentres_.setSize(m->n_faces);
    fm.faceAreas_.setSize(m->n_faces);
    for (int f = 0; f < m->n_faces; ++f)
    {
        face& fa = fm.faces_[f];
        for (int k = m->face_offsets[f]; k < m->face_offsets[f + 1]; ++k) fa.append(m->face_points[k]);
        fm.owner_[f] = m->owner[f];
        if (f < m->n_internal_faces) fm.neighbour_[f] = m->neighbour[f];
        fm.faceCentres_[f] = v3(om.Cf[f]);
        fm.faceAreas_[f] = v3(om.Sf[f]);
    }
    // primitiveMesh::calcCells: every face to its owner (ascending), then every internal face to its neighbour
    fm.cells_.setSize(m->n_cells);
    for (int f = 0; f < m->n_faces; ++f) fm.cells_[m->owner[f]].append(f);
    for (int f = 0; f < m->n_internal_faces; ++f) fm.cells_[m->neighbour[f]].append(f);
    fm.cellCentres_.setSize(m->n_cells);
    fm.cellVolumes_.setSize(m->n_cells);
    fm.cellPoints_.setSize(m->n_cells);
    fm.V_.setSize(m->n_cells);
    fm.pointCells_.setSize(m->n_points);
    for (int c = 0; c < m->n_cells; ++c)
    {
        fm.cellCentres_[c] = v3(om.C[c]);
        fm.cellVolumes_[c] = om.V[c];
        fm.V_[c] = om.V[c];
        fm.cellPoints_[c] = fm.cells_[c].labels(fm.faces_);
        for (label k = 0; k < fm.cellPoints_[c].size(); ++k) fm.pointCells_[fm.cellPoints_[c][k]].append(c);
    }
    fm.boundaryMesh_.patchID_.setSize(m->n_faces - m->n_internal_faces);
    for (int p = 0; p < m->n_patches; ++p)
    {
        const svof_patch& pt = m->patches[p];
        if (pt.kind == SVOF_PATCH_PROCESSOR && pt.size > 0) throw std::invalid_argument("processor patches: serial reference only");
        if (pt.kind == SVOF_PATCH_EMPTY) fm.boundaryMesh_.patches_.push_back(std::make_shared<emptyPolyPatch>(pt.start, pt.size));
        else fm.boundaryMesh_.patches_.push_back(std::make_shared<polyPatch>(pt.start, pt.size));
        for (int k = 0; k < pt.size; ++k) fm.boundaryMesh_.patchID_[pt.start + k - m->n_internal_faces] = p;
    }
    for (int d = 0; d < 3; ++d) fm.geometricD_[d] = geomD[d];
    fm.magSf_.reset(new surfaceScalarField(IOobject("magSf", "0", fm), fm, dimensionedScalar(dimless, 0.0)));
    setSurface(*fm.magSf_, fm, om.magSf.data());
    fm.Cf_.reset(new surfaceVectorField(IOobject("Cf", "0", fm), fm, dimensionedVector(dimless, vector())));
    setSurfaceVec(*fm.Cf_, fm, om.Cf);
    fm.Sf_.reset(new surfaceVectorField(IOobject("Sf", "0", fm), fm, dimensionedVector(dimless, vector())));
    setSurfaceVec(*fm.Sf_, fm, om.Sf);
    // volVectorField C: cell centres, patch values = face centres
    fm.C_.reset(new volVectorField(IOobject("C", "0", fm), fm, dimensionedVector(dimless, vector())));
    for (int c = 0; c < m->n_cells; ++c) (*fm.C_)[c] = v3(om.C[c]);
    for (label p = 0; p < fm.C_->boundaryField().size(); ++p)
    {
        const label start = fm.boundaryMesh()[p].start();
        for (label i = 0; i < fm.C_->boundaryField()[p].size(); ++i) fm.C_->boundaryFieldRef()[p][i] = v3(om.Cf[start + i]);
    }
}
}  // namespace

extern "C" {

// mesh_kind: 0 = dynamicFvMesh, 1 = dynamicRefineFvMesh (mapAlphaField acts), 2 = dynamicOversetFvMesh (cellTypes filter)
void* ref_solver_create(const svof_mesh* m, const svof_params* prm, int32_t mesh_kind)
{
    if (!m || !prm) return nullptr;
    Ref* r = new (std::nothrow) Ref;
    if (!r) return nullptr;
    try
    {
        r->S.init(*m, *prm);
        const ora::Mesh& om = r->S.mesh;
        if (mesh_kind == 1) r->fmp.reset(new dynamicRefineFvMesh);
        else if (mesh_kind == 2) r->fmp.reset(new dynamicOversetFvMesh);
        else r->fmp.reset(new dynamicFvMesh);
        fvMesh& fm = *r->fmp;
        fillMesh(fm, m, om, r->S.geomD);

        // fvSolution solvers."alpha.*"
        dictionary& d = fm.solverDict_;
        d.set("mixedCellTol", prm->mixed_cell_tol);
        // writePlicFields only adds the zeta copy of the flatness (reconstruction.C:444-472), which indexes the patch fields of
        // EMPTY patches as well (size 0 in an fvMesh): switched off on 2-D meshes, where the reference would write out of bounds
        bool hasEmpty = false;
        for (int p = 0; p < m->n_patches; ++p) hasEmpty = hasEmpty || (m->patches[p].kind == SVOF_PATCH_EMPTY && m->patches[p].size > 0);
        d.set("writePlicFields", (prm->write_plic_fields && !hasEmpty) ? 1.0 : 0.0);
        d.set("splitWarpedFace", prm->split_warped_face ? 1.0 : 0.0);
        d.setWord("orientationMethod", prm->orientation_method == SVOF_ORIENT_ALPHA_GRAD ? "alphaGrad"
                                     : prm->orientation_method == SVOF_ORIENT_ISO_RDF ? "isoRDF" : "isoAlphaGrad");
        d.set("tol", prm->rdf_tol);
        d.set("relTol", prm->rdf_rel_tol);
        d.set("iterations", prm->rdf_iterations);
        d.set("mapAlphaField", prm->map_alpha_field ? 1.0 : 0.0);
        d.set("nAlphaBounds", prm->n_alpha_bounds);
        d.set("snapTol", prm->snap_tol);
        d.set("clip", prm->clip ? 1.0 : 0.0);

        r->alpha = new volScalarField(IOobject("alpha.water", "0", fm), fm, dimensionedScalar(dimless, 0.0));
        r->phi = new surfaceScalarField(IOobject("phi", "0", fm), fm, dimensionedScalar(dimless, 0.0));
        r->U = new volVectorField(IOobject("U", "0", fm), fm, dimensionedVector(dimless, vector()));
        fm.objects_.push_back(std::make_pair(word("U"), static_cast<const void*>(r->U)));
        Ref* rp = r;
        // correctBoundaryConditions() of alpha: zeroGradient -> patchInternalField, fixedValue -> the value,
        // inletOutlet -> valueFraction = 1 - pos0(phi_p), value = f*inletValue + (1 - f)*patchInternalField
        r->alpha->bcHook = [rp](volScalarField& a)
        {
            const ora::Mesh& mesh = rp->S.mesh;
            for (label p = 0; p < a.boundaryField().size(); ++p)
            {
                const svof_patch& pt = mesh.patches[p];
                fvPatchField<scalar>& pf = a.boundaryFieldRef()[p];
                const fvsPatchField<scalar>& phip = rp->phi->boundaryField()[p];
                for (label i = 0; i < pf.size(); ++i)
                {
                    const scalar internal = a[mesh.owner[pt.start + i]];
                    if (pt.alpha_bc == SVOF_BC_FIXED_VALUE) pf[i] = pt.alpha_value;
                    else if (pt.alpha_bc == SVOF_BC_INLET_OUTLET)
                    {
                        const scalar vf = 1.0 - pos0(phip[i]);
                        pf[i] = vf * pt.alpha_value + (1.0 - vf) * internal;
                    }
                    else pf[i] = internal;
                }
            }
        };
        ora::Solver* Sp_ = &r->S;
        stubInterpolateCellPoint = [Sp_](const vector& pos, label celli)
        {
            const ora::vec v = Sp_->interpolateU(ora::vec(pos.x(), pos.y(), pos.z()), celli);
            return vector(v.x, v.y, v.z);
        };
        stubInfoStream().str("");
        stubInfoStream().precision(17);
        r->solver = new RefSolver(*r->alpha, *r->phi, *r->U);
        r->ctorLog = stubInfoStream().str();
        r->Sp.setSize(m->n_cells);
        r->Su.setSize(m->n_cells);
    }
    catch (const std::exception& e)
    {
        delete r;
        return nullptr;
    }
    return r;
}

void ref_solver_destroy(void* h) { delete static_cast<Ref*>(h); }

// the hooks are process-wide: (re)point them at this handle before every call into the reference
static void bindHooks(Ref* r)
{
    ora::Solver* Sp_ = &r->S;
    stubGradScheme = r->S.prm.alpha_grad_scheme;
    stubInterpolateCellPoint = [Sp_](const vector& pos, label celli)
    {
        const ora::vec v = Sp_->interpolateU(ora::vec(pos.x(), pos.y(), pos.z()), celli);
        return vector(v.x, v.y, v.z);
    };
    stubCpcStencil = [Sp_](const fvMesh&, label celli, labelList& st)
    {
        std::vector<ora::label> s;
        Sp_->cpcStencil(celli, s);
        st.setSize(label(s.size()));
        for (size_t i = 0; i < s.size(); ++i) st[label(i)] = s[i];
    };
    stubInfoStream().str("");
    stubInfoStream().precision(17);
}

// reconstruction::faceFlatness() [nF] as the reference's constructor computed it (updateFaceFlatness)
int ref_solver_face_flatness(void* h, double* out)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    const scalarField& z = r->solver->reconstructor().faceFlatness();
    for (label f = 0; f < z.size(); ++f) out[f] = z[f];
    return 0;
}

// alpha [nC] (patch values evaluated), phi [nF], U [3 nC], Ub [3 nBF]
int ref_solver_set_state(void* h, const double* alpha, const double* phi, const double* U, const double* Ub)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    fvMesh& fm = *r->fmp;
    const label nC = fm.nCells(), nF = fm.nFaces(), nIF = fm.nInternalFaces();
    if (phi) setSurface(*r->phi, fm, phi);
    if (alpha)
    {
        for (label c = 0; c < nC; ++c) (*r->alpha)[c] = alpha[c];
        r->alpha->correctBoundaryConditions();
    }
    if (U)
        for (label c = 0; c < nC; ++c)
        {
            (*r->U)[c] = v3(U, c);
            r->S.U[c] = ora::vec(U[3 * c], U[3 * c + 1], U[3 * c + 2]);
        }
    if (Ub)
        for (label b = 0; b < nF - nIF; ++b) r->S.Ub[b] = ora::vec(Ub[3 * b], Ub[3 * b + 1], Ub[3 * b + 2]);
    return 0;
}

// cellTypes [nC] of the overset stencil (cellCellStencil::cellType: 0 CALCULATED, 1 INTERPOLATED, 2 HOLE)
int ref_solver_set_cell_types(void* h, const int32_t* types)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    fvMesh& fm = *r->fmp;
    cellCellStencilObject& o = stubOversetStencil(fm);
    o.cellTypes_.setSize(fm.nCells());
    for (label c = 0; c < fm.nCells(); ++c) o.cellTypes_[c] = types[c];
    // cellMask as dynamicOversetFvMesh keeps it: 0 in holes, 1 elsewhere; zeroGradient patches
    if (!r->cellMask)
    {
        r->cellMask = new volScalarField(IOobject("cellMask", "0", fm), fm, dimensionedScalar(dimless, 1.0));
        fm.objects_.push_back(std::make_pair(word("cellMask"), static_cast<const void*>(r->cellMask)));
    }
    for (label c = 0; c < fm.nCells(); ++c) (*r->cellMask)[c] = (types[c] == cellCellStencil::HOLE) ? 0.0 : 1.0;
    for (label p = 0; p < r->cellMask->boundaryField().size(); ++p)
    {
        const label start = fm.boundaryMesh()[p].start();
        for (label i = 0; i < r->cellMask->boundaryField()[p].size(); ++i)
            r->cellMask->boundaryFieldRef()[p][i] = (*r->cellMask)[fm.faceOwner()[start + i]];
    }
    return 0;
}

int ref_solver_reconstruct(void* h, int32_t* nMixed)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    try
    {
        bindHooks(r);
        r->solver->reconstruct();
        r->log = stubInfoStream().str();
        if (nMixed) *nMixed = r->solver->reconstructor().mixedCells().size();
    }
    catch (const std::exception& e)
    {
        r->err = e.what();
        return -2;
    }
    return 0;
}

// mixedCells / cellStatus [nMixed], interfaceN / C / S [3 nC], interfaceD [nC] (any may be NULL)
int ref_solver_get_recon(void* h, int32_t* mixedCells, int32_t* cellStatus, double* iN, double* iD, double* iC, double* iS)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    const geometricVofExt::SimPLIC::reconstruction& rec = r->solver->reconstructor();
    const label nC = r->fmp->nCells();
    for (label i = 0; i < rec.mixedCells().size(); ++i)
    {
        if (mixedCells) mixedCells[i] = rec.mixedCells()[i];
        if (cellStatus) cellStatus[i] = rec.cellStatus()[i];
    }
    for (label c = 0; c < nC; ++c)
    {
        for (int d = 0; d < 3; ++d)
        {
            if (iN) iN[3 * c + d] = rec.interfaceN()[c][d];
            if (iC) iC[3 * c + d] = rec.interfaceC()[c][d];
            if (iS) iS[3 * c + d] = rec.interfaceS()[c][d];
        }
        if (iD) iD[c] = rec.interfaceD()[c];
    }
    return 0;
}

// runTime++ (alpha.oldTime() := alpha) then solveVofEqu::advect(Sp, Su); Sp / Su [nC] or NULL (= zeroField)
int ref_solver_advect(void* h, double dt, const double* Sp, const double* Su)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    try
    {
        bindHooks(r);
        fvMesh& fm = *r->fmp;
        const label nC = fm.nCells();
        fm.time_.setDeltaT(dt);
        r->alpha->storeOldTime();
        if (Sp) for (label c = 0; c < nC; ++c) r->Sp[c] = Sp[c];
        if (Su) for (label c = 0; c < nC; ++c) r->Su[c] = Su[c];
        if (Sp && Su) r->solver->advect(r->Sp, r->Su);
        else if (Sp) r->solver->advect(r->Sp, zeroField());
        else if (Su) r->solver->advect(zeroField(), r->Su);
        else r->solver->advect(zeroField(), zeroField());
        r->log = stubInfoStream().str();
    }
    catch (const std::exception& e)
    {
        r->err = e.what();
        return -2;
    }
    return 0;
}

// alpha [nC], alphaPhi [nF], alpha patch values [nBF] (any may be NULL)
int ref_solver_get_fields(void* h, double* alpha, double* alphaPhi, double* alphaB)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    fvMesh& fm = *r->fmp;
    const label nC = fm.nCells(), nF = fm.nFaces(), nIF = fm.nInternalFaces();
    if (alpha) for (label c = 0; c < nC; ++c) alpha[c] = (*r->alpha)[c];
    if (alphaPhi) getSurface(r->solver->alphaPhi(), fm, alphaPhi);
    if (alphaB)
    {
        for (label b = 0; b < nF - nIF; ++b) alphaB[b] = 0.0;
        for (label p = 0; p < r->alpha->boundaryField().size(); ++p)
        {
            const label start = fm.boundaryMesh()[p].start();
            for (label i = 0; i < r->alpha->boundaryField()[p].size(); ++i) alphaB[start + i - nIF] = r->alpha->boundaryField()[p][i];
        }
    }
    return 0;
}

// reconstruction::mapAlphaField() on a changing dynamicRefineFvMesh with dynamicMeshDict {lowerRefineLevel, upperRefineLevel}
int ref_solver_map_alpha(void* h, double lowerRefineLevel, double upperRefineLevel)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    try
    {
        bindHooks(r);
        stubDynamicMeshDict() = dictionary();
        stubDynamicMeshDict().set("lowerRefineLevel", lowerRefineLevel);
        stubDynamicMeshDict().set("upperRefineLevel", upperRefineLevel);
        r->fmp->changing_ = true;
        try { r->solver->mapAlphaField(); }
        catch (...) { r->fmp->changing_ = false; throw; }
        r->fmp->changing_ = false;
    }
    catch (const std::exception& e)
    {
        r->err = e.what();
        return -2;
    }
    return 0;
}

// which = 0: reconstruction::interface(), 1: reconstruction::subCellFaces(); sizes out, data via ref_solver_surface_copy
int ref_solver_surface(void* h, int32_t which, int32_t* nPoints, int32_t* nFaces, int32_t* nFacePoints, int32_t* nCells)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    try
    {
        bindHooks(r);
        pointField pts;
        faceList faces;
        labelList cells;
        if (which == 0)
        {
            geometricVofExt::SimPLIC::plicSurface s(r->solver->reconstructor().interface());
            pts = s.points_; faces = s.faces_; cells = s.meshCells_;
        }
        else
        {
            geometricVofExt::SimPLIC::reconstructedSubcellFaces s(r->solver->reconstructor().subCellFaces());
            pts = s.points_; faces = s.faces_; cells = s.meshCells_;
        }
        r->sPts.resize(3 * size_t(pts.size()));
        for (label i = 0; i < pts.size(); ++i)
            for (int d = 0; d < 3; ++d) r->sPts[3 * size_t(i) + d] = pts[i][d];
        r->sOff.assign(1, 0);
        r->sFacePts.clear();
        for (label f = 0; f < faces.size(); ++f)
        {
            for (label k = 0; k < faces[f].size(); ++k) r->sFacePts.push_back(faces[f][k]);
            r->sOff.push_back(int32_t(r->sFacePts.size()));
        }
        r->sCells.resize(cells.size());
        for (label i = 0; i < cells.size(); ++i) r->sCells[i] = cells[i];
        *nPoints = pts.size();
        *nFaces = faces.size();
        *nFacePoints = int32_t(r->sFacePts.size());
        *nCells = cells.size();
    }
    catch (const std::exception& e)
    {
        r->err = e.what();
        return -2;
    }
    return 0;
}
int ref_solver_surface_copy(void* h, double* points, int32_t* faceOffsets, int32_t* facePoints, int32_t* meshCells)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return -1;
    std::copy(r->sPts.begin(), r->sPts.end(), points);
    std::copy(r->sOff.begin(), r->sOff.end(), faceOffsets);
    std::copy(r->sFacePts.begin(), r->sFacePts.end(), facePoints);
    std::copy(r->sCells.begin(), r->sCells.end(), meshCells);
    return 0;
}

// Info<< text of the last reconstruct / advect ("SimPLIC::reconstruction: Number of mixed cells = ...", the bounding
// lines; 17 digits) and of the constructor ("SimPLIC::Mesh face flatness: min/max/avg = ...")
const char* ref_solver_log(void* h) { return h ? static_cast<Ref*>(h)->log.c_str() : ""; }
const char* ref_solver_ctor_log(void* h) { return h ? static_cast<Ref*>(h)->ctorLog.c_str() : ""; }
const char* ref_solver_error(void* h) { return h ? static_cast<Ref*>(h)->err.c_str() : "null handle"; }

}  // extern "C"
