// ref_advect.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE's own advection class.  oracle/build.py:build_ref_advect compiles this file
// together with the reference files, read where they lie,
//     /root/reference/src/SimPLIC/advection/advection.{H,C}, advectionTemplates.C
//     /root/reference/src/SimPLIC/cut/cutFace/cutFace.{H,C}          (cutCell.H is only parsed)
// against oracle/of_stub_adv/ (a stand-in for the OpenFOAM types they use; OpenFOAM itself is not installable here)
// into oracle/_ref/libref_advect.so.  Nothing of the reference is copied into this repository.
//
//   ref_advect_step  <->  svof_advect(dt, Sp, Su)      (advection::advect, advectionTemplates.C:352-418)
//
// The reconstruction the class reads (mixed-cell list, cut status, interfaceN/D/C) is an INPUT -- the reference's
// advection only holds a const reference to it -- and so are the OpenFOAM quantities outside the compiled files: mesh
// geometry and face flatness (taken from the oracle's mesh services), the patch-field evaluation of alpha
// (zeroGradient / fixedValue / inletOutlet, written here as the oracle writes it) and interpolationCellPoint (answered
// by the oracle's restatement).  What the tests then compare bitwise is everything advection.C / advectionTemplates.C
// themselves do: downwind-face selection, time-integrated face fluxes, the alpha update, the bounding sweeps, snap/clip,
// alphaPhi.
#include <cstdint>
#include <new>
#include <string>

#include "ora_solver.hpp"   // before the stand-in: it defines the macros Info / forAll

#define NoRepository
#include "advection.H"

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
}  // namespace Foam

namespace
{
using namespace Foam;
typedef geometricVofExt::SimPLIC::advection RefAdvection;
typedef geometricVofExt::SimPLIC::reconstruction RefReconstruction;

struct RefAdvect
{
    ora::Solver S;   // mesh services (geometry, flatness, U interpolation); its own advect() is never called here
    dynamicFvMesh fm;
    dictionary dict;
    volScalarField* alpha;
    surfaceScalarField* phi;
    volVectorField* U;
    RefReconstruction* rec;
    RefAdvection* adv;
    DimensionedField<scalar> Sp, Su;
    std::string log, err;
    RefAdvect() : alpha(nullptr), phi(nullptr), U(nullptr), rec(nullptr), adv(nullptr) {}
    ~RefAdvect()
    {
        delete adv;
        delete rec;
        delete U;
        delete phi;
        delete alpha;
    }
};

inline vector v3(const double* p, int64_t i) { return vector(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline vector v3(const ora::vec& v) { return vector(v.x, v.y, v.z); }

// scatter a face-indexed array (internal faces, then boundary faces in mesh order) into a surface field
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
}  // namespace

extern "C" {

void* ref_advect_create(const svof_mesh* m, const svof_params* prm)
{
    if (!m || !prm) return nullptr;
    RefAdvect* r = new (std::nothrow) RefAdvect;
    if (!r) return nullptr;
    try
    {
        r->S.init(*m, *prm);
        const ora::Mesh& om = r->S.mesh;
        dynamicFvMesh& fm = r->fm;
        fm.nInternalFaces_ = m->n_internal_faces;
        fm.points_.setSize(m->n_points);
        for (int i = 0; i < m->n_points; ++i) fm.points_[i] = v3(m->points, i);
        fm.faces_.setSize(m->n_faces);
        fm.owner_.setSize(m->n_faces);
        fm.neighbour_.setSize(m->n_internal_faces);
        fm.faceCentres_.setSize(m->n_faces);
        for (int f = 0; f < m->n_faces; ++f)
        {
            face& fa = fm.faces_[f];
            for (int k = m->face_offsets[f]; k < m->face_offsets[f + 1]; ++k) fa.append(m->face_points[k]);
            fm.owner_[f] = m->owner[f];
            if (f < m->n_internal_faces) fm.neighbour_[f] = m->neighbour[f];
            fm.faceCentres_[f] = v3(om.Cf[f]);
        }
        // primitiveMesh::calcCells: every face to its owner (ascending), then every internal face to its neighbour
        fm.cells_.setSize(m->n_cells);
        for (int f = 0; f < m->n_faces; ++f) fm.cells_[m->owner[f]].append(f);
        for (int f = 0; f < m->n_internal_faces; ++f) fm.cells_[m->neighbour[f]].append(f);
        fm.cellCentres_.setSize(m->n_cells);
        fm.cellVolumes_.setSize(m->n_cells);
        fm.cellPoints_.setSize(m->n_cells);
        fm.V_.setSize(m->n_cells);
        for (int c = 0; c < m->n_cells; ++c)
        {
            fm.cellCentres_[c] = v3(om.C[c]);
            fm.cellVolumes_[c] = om.V[c];
            fm.V_[c] = om.V[c];
            fm.cellPoints_[c] = fm.cells_[c].labels(fm.faces_);
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
        fm.magSf_.reset(new surfaceScalarField(IOobject("magSf", "0", fm), fm, dimensionedScalar(dimless, 0.0)));
        setSurface(*fm.magSf_, fm, om.magSf.data());
        fm.Cf_.reset(new surfaceVectorField(IOobject("Cf", "0", fm), fm, dimensionedVector(dimless, vector())));
        for (int f = 0; f < m->n_internal_faces; ++f) (*fm.Cf_)[f] = v3(om.Cf[f]);
        fm.C_.reset(new volVectorField(IOobject("C", "0", fm), fm, dimensionedVector(dimless, vector())));
        for (int c = 0; c < m->n_cells; ++c) (*fm.C_)[c] = v3(om.C[c]);

        r->alpha = new volScalarField(IOobject("alpha.water", "0", fm), fm, dimensionedScalar(dimless, 0.0));
        r->phi = new surfaceScalarField(IOobject("phi", "0", fm), fm, dimensionedScalar(dimless, 0.0));
        r->U = new volVectorField(IOobject("U", "0", fm), fm, dimensionedVector(dimless, vector()));
        // correctBoundaryConditions() of alpha: zeroGradient -> patchInternalField, fixedValue -> the value,
        // inletOutlet -> valueFraction = 1 - pos0(phi_p), value = f*inletValue + (1 - f)*patchInternalField
        RefAdvect* rp = r;
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
        r->rec = new RefReconstruction(fm);
        r->rec->faceFlatness_.setSize(m->n_faces);
        for (int f = 0; f < m->n_faces; ++f) r->rec->faceFlatness_[f] = om.faceFlatness[f];
        r->dict.set("nAlphaBounds", prm->n_alpha_bounds);
        r->dict.set("snapTol", prm->snap_tol);
        r->dict.set("clip", prm->clip ? 1.0 : 0.0);
        r->adv = new RefAdvection(*r->alpha, *r->phi, *r->U, *r->rec, r->dict);
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

void ref_advect_destroy(void* h) { delete static_cast<RefAdvect*>(h); }

// One advection::advect(Sp, Su) of the reference class on the given state.
//   alpha [nC], phi [nF], U [3 nC], Ub [3 nBF]; reconstruction: mixedCells / cellStatus [nMixed], iN / iC [3 nC], iD [nC]
//   Sp, Su [nC] or NULL (= zeroField, as plicVofAdvectionFoam passes)
//   out: alpha_out [nC], alphaPhi_out [nF], alphaB_out [nBF]
int ref_advect_step(void* h, const double* alpha, const double* phi, const double* U, const double* Ub, int32_t nMixed,
                    const int32_t* mixedCells, const int32_t* cellStatus, const double* iN, const double* iD, const double* iC,
                    double dt, const double* Sp, const double* Su, double* alpha_out, double* alphaPhi_out, double* alphaB_out)
{
    RefAdvect* r = static_cast<RefAdvect*>(h);
    if (!r) return -1;
    try
    {
        dynamicFvMesh& fm = r->fm;
        const label nC = fm.nCells(), nF = fm.nFaces(), nIF = fm.nInternalFaces();
        fm.time_.setDeltaT(dt);
        setSurface(*r->phi, fm, phi);
        for (label c = 0; c < nC; ++c) (*r->alpha)[c] = alpha[c];
        r->alpha->correctBoundaryConditions();
        r->alpha->storeOldTime();   // runTime++ of the time loop
        for (label c = 0; c < nC; ++c)
        {
            (*r->U)[c] = v3(U, c);
            r->S.U[c] = ora::vec(U[3 * c], U[3 * c + 1], U[3 * c + 2]);
        }
        for (label b = 0; b < nF - nIF; ++b) r->S.Ub[b] = ora::vec(Ub[3 * b], Ub[3 * b + 1], Ub[3 * b + 2]);
        ora::Solver* Sp_ = &r->S;
        stubInterpolateCellPoint = [Sp_](const vector& pos, label celli)
        {
            const ora::vec v = Sp_->interpolateU(ora::vec(pos.x(), pos.y(), pos.z()), celli);
            return vector(v.x, v.y, v.z);
        };
        r->rec->mixedCells_.clear();
        r->rec->cellStatus_.clear();
        for (int i = 0; i < nMixed; ++i)
        {
            r->rec->mixedCells_.append(mixedCells[i]);
            r->rec->cellStatus_.append(cellStatus[i]);
        }
        for (label c = 0; c < nC; ++c)
        {
            r->rec->interfaceN_[c] = v3(iN, c);
            r->rec->interfaceC_[c] = v3(iC, c);
            r->rec->interfaceD_[c] = iD[c];
        }
        stubInfoStream().str("");
        stubInfoStream().precision(17);
        if (Sp) for (label c = 0; c < nC; ++c) r->Sp[c] = Sp[c];
        if (Su) for (label c = 0; c < nC; ++c) r->Su[c] = Su[c];
        if (Sp && Su) r->adv->advect(r->Sp, r->Su);
        else if (Sp) r->adv->advect(r->Sp, zeroField());
        else if (Su) r->adv->advect(zeroField(), r->Su);
        else r->adv->advect(zeroField(), zeroField());
        r->log = stubInfoStream().str();
        for (label c = 0; c < nC; ++c) alpha_out[c] = (*r->alpha)[c];
        getSurface(r->adv->alphaPhi(), fm, alphaPhi_out);
        for (label b = 0; b < nF - nIF; ++b) alphaB_out[b] = 0.0;
        for (label p = 0; p < r->alpha->boundaryField().size(); ++p)
        {
            const label start = fm.boundaryMesh()[p].start();
            for (label i = 0; i < r->alpha->boundaryField()[p].size(); ++i) alphaB_out[start + i - nIF] = r->alpha->boundaryField()[p][i];
        }
    }
    catch (const std::exception& e)
    {
        r->err = e.what();
        return -2;
    }
    return 0;
}

// the Info<< lines of the last step ("SimPLIC::advection: Before / After conservative bounding: ...", 17 digits)
const char* ref_advect_log(void* h) { return h ? static_cast<RefAdvect*>(h)->log.c_str() : ""; }
const char* ref_advect_error(void* h) { return h ? static_cast<RefAdvect*>(h)->err.c_str() : "null handle"; }

}  // extern "C"
