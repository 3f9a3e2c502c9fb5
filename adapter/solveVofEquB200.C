/*---------------------------------------------------------------------------*\
  solveVofEquB200.C -- see solveVofEquB200.H.  Everything numerical happens behind include/svof.h; this file marshals
  OpenFOAM fields into the flat arrays of that ABI and back, and mirrors the reference's log lines.
\*---------------------------------------------------------------------------*/
#include "solveVofEquB200.H"

#include <cstring>
#include <sstream>
#include <vector>

namespace Foam
{
namespace geometricVofExt
{
namespace SimPLIC
{

const char* const solveVofEqu::typeName = "solveVofEqu";

// * * * * * * * * * * * * * * * * helpers  * * * * * * * * * * * * * * * * * //

void solveVofEqu::check(int rc, const svof_handle* h, const char* what)
{
    if (rc < 0)
    {
        // the reference aborts the same way (reconstruction.C:610-625, advectionTemplates.C:58-63)
        FatalErrorInFunction
            << what << ": svof error " << rc << ": " << svof_last_error(h)
            << abort(FatalError);
    }
}

//- Text of the first token of a dictionary entry ("nAlphaBounds 3;" -> "3")
static std::string entryText(const dictionary& dict, const word& key)
{
#ifdef OPENFOAM_STUB_H
    return dict.firstTokenText(key);
#else
    ITstream& is = dict.lookup(key);
    OStringStream os;
    os << token(is);
    return os.str();
#endif
}

//- fvSolution solvers."alpha.*" -> svof_params (reconstruction.C:502-516, advection.C:455-457)
void solveVofEqu::readControls(svof_params& p) const
{
    svof_params_default(&p);
    const List<word> keys(dict_.toc());
    forAll(keys, i)
    {
        const int rc = svof_params_set(&p, keys[i].c_str(), entryText(dict_, keys[i]).c_str());
        if (rc == SVOF_ERR_BAD_CONFIG)
        {
            FatalErrorInFunction
                << "Orientation vector calculation method '" << entryText(dict_, keys[i]) << "' is not valid. "
                << "Valid methods are (alphaGrad isoAlphaGrad isoRDF)" << abort(FatalError);   // reconstruction.C:612-624
        }
        // SVOF_ERR_INVALID_ARG: a key of the same dictionary that belongs to someone else (solver tolerances ...)
    }
#ifndef OPENFOAM_STUB_H
    // orientationMethod alphaGrad evaluates fvc::grad(alpha1_, "grad(alpha1)") (reconstruction.C:78) with the CASE's
    // gradient scheme: hand the fvSchemes entry over ("Gauss linear" | "Gauss pointLinear")
    if (p.orientation_method == SVOF_ORIENT_ALPHA_GRAD)
    {
        ITstream& is = mesh_.gradScheme("grad(alpha1)");
        std::string text;
        while (!is.eof())
        {
            OStringStream os;
            os << token(is);
            if (os.str().empty()) break;
            text += (text.empty() ? "" : " ") + os.str();
        }
        if (svof_params_set(&p, "gradSchemes", text.c_str()) != SVOF_OK)
        {
            FatalErrorInFunction
                << "gradSchemes entry for grad(alpha1) is '" << text << "'; the device path has Gauss linear and Gauss pointLinear"
                << abort(FatalError);
        }
    }
#endif
}

//- polyBoundaryMesh + alpha boundary conditions -> svof_patch table
static List<svof_patch> patchTable(const fvMesh& mesh, const volScalarField& alpha1)
{
    const polyBoundaryMesh& pbm = mesh.boundaryMesh();
    svof_patch zero;
    std::memset(&zero, 0, sizeof(zero));
    List<svof_patch> patches(pbm.size(), zero);
    forAll(pbm, pi)
    {
        svof_patch& q = patches[pi];
        q.start = pbm[pi].start();
        q.size = pbm[pi].size();
        q.nbr_rank = -1;
        q.kind = (pbm[pi].type() == "empty") ? SVOF_PATCH_EMPTY
               : (pbm[pi].type() == "processor") ? SVOF_PATCH_PROCESSOR : SVOF_PATCH_GENERIC;
        if (q.kind == SVOF_PATCH_PROCESSOR) q.nbr_rank = pbm[pi].neighbProcNo();
        const fvPatchScalarField& bf = alpha1.boundaryField()[pi];
        q.alpha_bc = SVOF_BC_ZERO_GRADIENT;
        q.alpha_value = 0;
        if (bf.type() == "inletOutlet")
        {
            q.alpha_bc = SVOF_BC_INLET_OUTLET;
            const scalarField& rv = static_cast<const inletOutletFvPatchScalarField&>(bf).refValue();
            q.alpha_value = rv.size() ? rv[0] : 0;
        }
        else if (bf.type() == "fixedValue" && bf.size())
        {
            q.alpha_bc = SVOF_BC_FIXED_VALUE;
            q.alpha_value = bf[0];
        }
    }
    return patches;
}

//- faceList -> CSR
static void flattenFaces(const faceList& fs, labelList& off, labelList& pts)
{
    off.setSize(fs.size() + 1, 0);
    off[0] = 0;
    forAll(fs, i) off[i + 1] = off[i] + fs[i].size();
    pts.setSize(off.last());
    forAll(fs, i)
    {
        forAll(fs[i], k) pts[off[i] + k] = fs[i][k];
    }
}

// * * * * * * * * * * * * * * * * constructors * * * * * * * * * * * * * * * //

solveVofEqu::solveVofEqu(volScalarField& alpha1, const surfaceScalarField& phi, const volVectorField& U)
:
    mesh_(alpha1.mesh()),
    alpha1_(alpha1),
    phi_(phi),
    U_(U),
    dict_(mesh_.solverDict(alpha1.name())),          // solveVofEqu.C:68
    h_(nullptr),
    sub_(nullptr),
    nCellsDev_(0),
    nFacesDev_(0),
    nInternalDev_(0),
    alphaOnDevice_(false),
    alphaPhi_("alphaPhi", mesh_, 0.0),
    mapAlphaFieldOn_(false),
    reconstructor_(alpha1, *this)
{
    svof_params p;
    readControls(p);
    mapAlphaFieldOn_ = p.map_alpha_field != 0;
    if (Pstream::parRun()) createParallel(p); else createSerial(p);
    double fmin, fmax, favg;
    svof_get_info(h_, SVOF_I_FLATNESS_MIN, &fmin);
    svof_get_info(h_, SVOF_I_FLATNESS_MAX, &fmax);
    svof_get_info(h_, SVOF_I_FLATNESS_AVG, &favg);
    Info<< "SimPLIC::Mesh face flatness: min/max/avg = " << fmin << "/" << fmax << "/" << favg << endl;   // reconstruction.C:442-447
}

solveVofEqu::~solveVofEqu()
{
    svof_destroy(h_);
    if (sub_) svof_submesh_free(sub_);
}

//- One rank, one GPU, the whole mesh: OpenFOAM's own storage is handed over as it is (label == int32)
void solveVofEqu::createSerial(const svof_params& p)
{
    labelList faceOff, facePts;
    flattenFaces(mesh_.faces(), faceOff, facePts);
    List<svof_patch> patches(patchTable(mesh_, alpha1_));
    svof_mesh m;
    std::memset(&m, 0, sizeof(m));
    m.n_points = mesh_.nPoints();
    m.n_faces = mesh_.nFaces();
    m.n_internal_faces = mesh_.nInternalFaces();
    m.n_cells = mesh_.nCells();
    m.n_patches = patches.size();
    m.points = reinterpret_cast<const double*>(mesh_.points().cdata());   // vector == 3 contiguous doubles
    m.face_offsets = faceOff.cdata();
    m.face_points = facePts.cdata();
    m.owner = mesh_.faceOwner().cdata();
    m.neighbour = mesh_.faceNeighbour().cdata();
    m.patches = patches.cdata();
    // OpenFOAM's own geometry, so both sides use identical Cf/Sf/C/V
    m.Cf = reinterpret_cast<const double*>(mesh_.faceCentres().cdata());
    m.Sf = reinterpret_cast<const double*>(mesh_.faceAreas().cdata());
    m.C = reinterpret_cast<const double*>(mesh_.cellCentres().cdata());
    m.V = mesh_.cellVolumes().cdata();
    svof_comm c;
    std::memset(&c, 0, sizeof(c));
    c.rank = 0;
    c.world_size = 1;
    c.device = -1;
    check(svof_create(&m, &p, &c, &h_), nullptr, "svof_create");
    nCellsDev_ = m.n_cells;
    nFacesDev_ = m.n_faces;
    nInternalDev_ = m.n_internal_faces;
}

//- mpirun: this rank's cells + ghost layers cut out of the UNDECOMPOSED mesh (present in every decomposed case,
//  <case>/constant/polyMesh) with processorN/constant/polyMesh/{cell,face}ProcAddressing giving the ownership; the
//  processor patches themselves are not used (include/svof.h "decomposed runs").
void solveVofEqu::createParallel(const svof_params& p)
{
    const label me = Pstream::myProcNo(), nProcs = Pstream::nProcs();
#ifdef OPENFOAM_STUB_H
    const polyMesh& gMesh = *stubGlobalCase::mesh();
    auto addressing = [](label proc) -> const labelList& { return *stubGlobalCase::cellProcAddressing(proc); };
#else
    Time gTime(mesh_.time().rootPath(), mesh_.time().globalCaseName());
    polyMesh gMeshObj(IOobject(polyMesh::defaultRegion, gTime.constant(), gTime, IOobject::MUST_READ));
    const polyMesh& gMesh = gMeshObj;
    PtrList<labelIOList> addr(nProcs);
    auto addressing = [&](label proc) -> const labelList&
    {
        if (!addr.set(proc))
            addr.set(proc, new labelIOList(IOobject("cellProcAddressing",
                     gTime.rootPath()/gTime.globalCaseName()/("processor" + Foam::name(proc))/"constant"/polyMesh::meshSubDir,
                     gTime, IOobject::MUST_READ)));
        return addr[proc];
    };
#endif
    labelList cellRank(gMesh.nCells(), -1);
    for (label proc = 0; proc < nProcs; ++proc)
    {
        const labelList& a = addressing(proc);
        forAll(a, i) cellRank[a[i]] = proc;
    }
    labelList faceOff, facePts;
    flattenFaces(gMesh.faces(), faceOff, facePts);
    // the global patch table: same patches as the local mesh minus its processor patches (decomposePar keeps the order)
    List<svof_patch> localPatches(patchTable(mesh_, alpha1_));
    List<svof_patch> patches;
    forAll(gMesh.boundaryMesh(), pi)
    {
        svof_patch q = localPatches[pi];
        q.start = gMesh.boundaryMesh()[pi].start();
        q.size = gMesh.boundaryMesh()[pi].size();
        patches.append(q);
    }
    svof_mesh g;
    std::memset(&g, 0, sizeof(g));
    g.n_points = gMesh.nPoints();
    g.n_faces = gMesh.nFaces();
    g.n_internal_faces = gMesh.nInternalFaces();
    g.n_cells = gMesh.nCells();
    g.n_patches = patches.size();
    g.points = reinterpret_cast<const double*>(gMesh.points().cdata());
    g.face_offsets = faceOff.cdata();
    g.face_points = facePts.cdata();
    g.owner = gMesh.faceOwner().cdata();
    g.neighbour = gMesh.faceNeighbour().cdata();
    g.patches = patches.cdata();
    const label layers = p.n_alpha_bounds + 2;      // dependency radius of one step (include/svof.h)
    if (svof_decompose(&g, cellRank.cdata(), me, layers, &sub_) < 0)
    {
        FatalErrorInFunction << "svof_decompose: " << svof_decomp_last_error() << abort(FatalError);
    }
    svof_mesh m;
    svof_submesh_mesh(sub_, &m);
    int32_t nOwned = 0;
    const int32_t *cellGlobal, *cellOwnerRank, *ownedLocal, *faceGlobal, *faceOwnerRank, *faceFlip;
    svof_submesh_maps(sub_, &nOwned, &cellGlobal, &cellOwnerRank, nullptr, &ownedLocal, &faceGlobal, nullptr);
    svof_submesh_face_maps(sub_, &faceOwnerRank, &faceFlip);
    if (nOwned != mesh_.nCells())
    {
        FatalErrorInFunction << "cellProcAddressing does not match this processor mesh" << abort(FatalError);
    }
    ownedLocal_.setSize(nOwned);
    for (label i = 0; i < nOwned; ++i) ownedLocal_[i] = ownedLocal[i];   // decomposePar keeps the global cell order
    // NCCL bootstrap: rank 0 draws the id, Pstream carries it
    List<char> id(128, 0);
    if (me == 0) check(svof_comm_unique_id(id.data()), nullptr, "svof_comm_unique_id");
    Pstream::broadcast(id);
    svof_comm c;
    std::memset(&c, 0, sizeof(c));
    c.rank = me;
    c.world_size = nProcs;
    c.device = -1;                                   // rank % visible devices
    c.nccl_unique_id = id.cdata();
    check(svof_create(&m, &p, &c, &h_), nullptr, "svof_create");
    check(svof_halo_setup(h_, cellGlobal, cellOwnerRank), h_, "svof_halo_setup");
    check(svof_halo_setup_faces(h_, faceGlobal, faceOwnerRank, faceFlip), h_, "svof_halo_setup_faces");
    nCellsDev_ = m.n_cells;
    nFacesDev_ = m.n_faces;
    nInternalDev_ = m.n_internal_faces;
    // OpenFOAM face f of this processor mesh -> sub-mesh face: through the global label (faceProcAddressing holds
    // global label + 1, negative when the processor face is flipped)
#ifdef OPENFOAM_STUB_H
    faceLocal_.setSize(0);                           // (the stub has no faceProcAddressing; serial tests only)
#else
    labelIOList faceAddr(IOobject("faceProcAddressing", mesh_.facesInstance(), polyMesh::meshSubDir, mesh_, IOobject::MUST_READ));
    Map<label> globalToSub(2*m.n_faces);
    for (label i = 0; i < m.n_faces; ++i) globalToSub.insert(faceGlobal[i], i);
    faceLocal_.setSize(mesh_.nFaces());
    forAll(faceLocal_, f)
    {
        const label gf = mag(faceAddr[f]) - 1;
        const label sf = globalToSub[gf];
        // sign: processor face flipped against global XOR sub-mesh face flipped against global
        const bool flipped = (faceAddr[f] < 0) != (faceFlip[sf] != 0);
        faceLocal_[f] = flipped ? -(sf + 1) : (sf + 1);
    }
#endif
}

// * * * * * * * * * * * * * * * * field traffic * * * * * * * * * * * * * * * //

void solveVofEqu::pushAlpha()
{
    if (!sub_)
    {
        check(svof_set_alpha(h_, alpha1_.primitiveField().cdata()), h_, "svof_set_alpha");
    }
    else
    {
        alphaFlat_.setSize(nCellsDev_, 0.0);
        forAll(ownedLocal_, i) alphaFlat_[ownedLocal_[i]] = alpha1_.primitiveField()[i];
        check(svof_set_alpha(h_, alphaFlat_.cdata()), h_, "svof_set_alpha");
        check(svof_halo_exchange(h_), h_, "svof_halo_exchange");        // ghost values from their owners
    }
    alphaOnDevice_ = true;
}

//- phi (internal + patches in order) and U / U boundary values in the flat layout of include/svof.h
void solveVofEqu::pushFluxes()
{
    const label nBF = nFacesDev_ - nInternalDev_;
    phiFlat_.setSize(nFacesDev_, 0.0);
    UbFlat_.setSize(3*nBF, 0.0);
    if (!sub_)
    {
        forAll(phi_.primitiveField(), f) phiFlat_[f] = phi_.primitiveField()[f];
        forAll(phi_.boundaryField(), pi)
        {
            const fvsPatchScalarField& pf = phi_.boundaryField()[pi];          // empty patches: size 0
            const label start = mesh_.boundaryMesh()[pi].start();
            forAll(pf, i) phiFlat_[start + i] = pf[i];
        }
        forAll(U_.boundaryField(), pi)
        {
            const fvPatchVectorField& pf = U_.boundaryField()[pi];
            const label off = mesh_.boundaryMesh()[pi].start() - mesh_.nInternalFaces();
            forAll(pf, i) for (direction d = 0; d < 3; ++d) UbFlat_[3*(off + i) + d] = pf[i][d];
        }
        check(svof_set_phi(h_, phiFlat_.cdata()), h_, "svof_set_phi");
        check(svof_set_U(h_, reinterpret_cast<const double*>(U_.primitiveField().cdata()), UbFlat_.cdata()), h_, "svof_set_U");
        return;
    }
    // parallel: scatter this rank's own values into the sub-mesh layout, the ghosts come from their owners
    auto put = [&](label f, scalar v)
    {
        const label s = faceLocal_[f];
        if (s > 0) phiFlat_[s - 1] = v; else phiFlat_[-s - 1] = -v;
    };
    forAll(phi_.primitiveField(), f) put(f, phi_.primitiveField()[f]);
    forAll(phi_.boundaryField(), pi)
    {
        const fvsPatchScalarField& pf = phi_.boundaryField()[pi];
        const label start = mesh_.boundaryMesh()[pi].start();
        forAll(pf, i) put(start + i, pf[i]);
    }
    UFlat_.setSize(3*nCellsDev_, 0.0);
    forAll(ownedLocal_, i) for (direction d = 0; d < 3; ++d) UFlat_[3*ownedLocal_[i] + d] = U_.primitiveField()[i][d];
    forAll(U_.boundaryField(), pi)
    {
        if (mesh_.boundaryMesh()[pi].type() == "processor") continue;
        const fvPatchVectorField& pf = U_.boundaryField()[pi];
        const label start = mesh_.boundaryMesh()[pi].start();
        forAll(pf, i)
        {
            const label s = mag(faceLocal_[start + i]) - 1 - nInternalDev_;
            for (direction d = 0; d < 3; ++d) UbFlat_[3*s + d] = pf[i][d];
        }
    }
    check(svof_set_phi(h_, phiFlat_.cdata()), h_, "svof_set_phi");
    check(svof_set_U(h_, UFlat_.cdata(), UbFlat_.cdata()), h_, "svof_set_U");
    check(svof_halo_exchange_inputs(h_), h_, "svof_halo_exchange_inputs");
}

//- alpha1 and alphaPhi back into OpenFOAM's fields
void solveVofEqu::pullResults()
{
    alphaFlat_.setSize(nCellsDev_);
    check(int(svof_get_field(h_, SVOF_F_ALPHA, alphaFlat_.data(), nCellsDev_) < 0 ? -1 : 0), h_, "svof_get_field(alpha)");
    scalarField& a = alpha1_.primitiveFieldRef();
    if (!sub_) { forAll(a, i) a[i] = alphaFlat_[i]; }
    else { forAll(a, i) a[i] = alphaFlat_[ownedLocal_[i]]; }
    alpha1_.correctBoundaryConditions();
    alphaPhiFlat_.setSize(nFacesDev_);
    check(int(svof_get_field(h_, SVOF_F_ALPHA_PHI, alphaPhiFlat_.data(), nFacesDev_) < 0 ? -1 : 0), h_, "svof_get_field(alphaPhi)");
    auto get = [&](label f) -> scalar
    {
        if (!sub_) return alphaPhiFlat_[f];
        const label s = faceLocal_[f];
        return s > 0 ? alphaPhiFlat_[s - 1] : -alphaPhiFlat_[-s - 1];
    };
    forAll(alphaPhi_.primitiveFieldRef(), f) alphaPhi_.primitiveFieldRef()[f] = get(f);
    forAll(alphaPhi_.boundaryFieldRef(), pi)
    {
        fvsPatchScalarField& pf = alphaPhi_.boundaryFieldRef()[pi];
        const label start = mesh_.boundaryMesh()[pi].start();
        forAll(pf, i) pf[i] = get(start + i);
    }
}

const double* solveVofEqu::sourcePtr(const volScalarField::Internal& f, scalarField& flat)
{
    if (!sub_) return f.field().cdata();
    flat.setSize(nCellsDev_, 0.0);
    forAll(ownedLocal_, i) flat[ownedLocal_[i]] = f.field()[i];
    return flat.cdata();
}

// * * * * * * * * * * * * * * * * member functions * * * * * * * * * * * * * //

void solveVofEqu::reconstruct()
{
    reconstructor_.reconstruct();
}

void solveVofEqu::setCellTypes(const labelList& cellTypes)
{
    if (cellTypes.size() == 0)
    {
        check(svof_set_cell_types(h_, nullptr), h_, "svof_set_cell_types");
        return;
    }
    if (sub_)
    {
        FatalErrorInFunction
            << "overset cell types on a decomposed device mesh need the ghost cells' types as well: not supported"
            << abort(FatalError);
    }
    if (cellTypes.size() != mesh_.nCells())
    {
        FatalErrorInFunction << "cellTypes has " << cellTypes.size() << " entries for " << mesh_.nCells() << " cells"
            << abort(FatalError);
    }
    std::vector<int32_t> t(cellTypes.size());
    forAll(cellTypes, c) t[c] = cellTypes[c];
    check(svof_set_cell_types(h_, t.data()), h_, "svof_set_cell_types");
}

void solveVofEqu::advectFlat(const double* Sp, const double* Su)
{
    if (!alphaOnDevice_) pushAlpha();
    pushFluxes();
    check(svof_advect(h_, mesh_.time().deltaTValue(), Sp, Su), h_, "svof_advect");
    double mn0, mx0, mn1, mx1;
    svof_get_info(h_, SVOF_I_MIN_ALPHA_BEFORE, &mn0);
    svof_get_info(h_, SVOF_I_MAX_ALPHA_M1_BEFORE, &mx0);
    svof_get_info(h_, SVOF_I_MIN_ALPHA_AFTER, &mn1);
    svof_get_info(h_, SVOF_I_MAX_ALPHA_M1_AFTER, &mx1);
    Info<< "SimPLIC::advection: Before conservative bounding: min(alpha) = "
        << mn0 << ", max(alpha) = 1 + " << mx0 << endl;                           // advectionTemplates.C:133
    Info<< "SimPLIC::advection: After  conservative bounding: min(alpha) = "
        << mn1 << ", max(alpha) = 1 + " << mx1 << endl;                           // advectionTemplates.C:211
    check(svof_synchronize(h_), h_, "svof_synchronize");                            // device capacity flags surface here
    pullResults();
    alphaOnDevice_ = false;     // the caller owns alpha1 (and alpha1.oldTime()) between calls: alphaEqnSubCycle.H:13-27
}

void solveVofEqu::mapAlphaField()
{
    if (!mesh_.changing() || !mapAlphaFieldOn_) return;       // reconstruction.C:727-732
    if (sub_)
    {
        FatalErrorInFunction << "mapAlphaField in a decomposed run: re-run svof_decompose on the refined global mesh (not wired)"
            << abort(FatalError);
    }
    // the mesh under us was refined: rebuild every mesh table of the handle, then hand over the fields OpenFOAM mapped
    labelList faceOff, facePts;
    flattenFaces(mesh_.faces(), faceOff, facePts);
    List<svof_patch> patches(patchTable(mesh_, alpha1_));
    svof_mesh m;
    std::memset(&m, 0, sizeof(m));
    m.n_points = mesh_.nPoints();
    m.n_faces = mesh_.nFaces();
    m.n_internal_faces = mesh_.nInternalFaces();
    m.n_cells = mesh_.nCells();
    m.n_patches = patches.size();
    m.points = reinterpret_cast<const double*>(mesh_.points().cdata());
    m.face_offsets = faceOff.cdata();
    m.face_points = facePts.cdata();
    m.owner = mesh_.faceOwner().cdata();
    m.neighbour = mesh_.faceNeighbour().cdata();
    m.patches = patches.cdata();
    m.Cf = reinterpret_cast<const double*>(mesh_.faceCentres().cdata());
    m.Sf = reinterpret_cast<const double*>(mesh_.faceAreas().cdata());
    m.C = reinterpret_cast<const double*>(mesh_.cellCentres().cdata());
    m.V = mesh_.cellVolumes().cdata();
    check(svof_update_mesh(h_, &m), h_, "svof_update_mesh");
    nCellsDev_ = m.n_cells;
    nFacesDev_ = m.n_faces;
    nInternalDev_ = m.n_internal_faces;
    if (label(interfaceNMapped_.size()) != mesh_.nCells() || label(interfaceDMapped_.size()) != mesh_.nCells())
    {
        FatalErrorInFunction << "interfaceN/interfaceD were not mapped onto the refined mesh" << abort(FatalError);
    }
    pushAlpha();
    check(svof_set_interface(h_, reinterpret_cast<const double*>(interfaceNMapped_.cdata()), interfaceDMapped_.cdata()), h_, "svof_set_interface");
    const dictionary& refineDict = mesh_.dynamicRefineCoeffs();             // dynamicMeshDict.dynamicRefineFvMeshCoeffs
    const scalar lower = std::stod(entryText(refineDict, "lowerRefineLevel"));
    const scalar upper = std::stod(entryText(refineDict, "upperRefineLevel"));
    check(svof_map_alpha_field(h_, lower, upper), h_, "svof_map_alpha_field");
    alphaFlat_.setSize(nCellsDev_);
    check(int(svof_get_field(h_, SVOF_F_ALPHA, alphaFlat_.data(), nCellsDev_) < 0 ? -1 : 0), h_, "svof_get_field(alpha)");
    scalarField& a = alpha1_.primitiveFieldRef();
    forAll(a, i) a[i] = alphaFlat_[i];
    alpha1_.correctBoundaryConditions();                                    // alpha1_.oldTime() = alpha1_ is the caller's storage
}

tmp<surfaceScalarField> solveVofEqu::getRhoPhi(const dimensionedScalar rho1, const dimensionedScalar rho2) const
{
    // advection.H:329-343: (rho1 - rho2)*alphaPhi + rho2*phi
    surfaceScalarField* r = new surfaceScalarField("rhoPhi", mesh_, 0.0);
    const scalar d = rho1.value() - rho2.value(), r2 = rho2.value();
    forAll(r->primitiveFieldRef(), f) r->primitiveFieldRef()[f] = d*alphaPhi_.primitiveField()[f] + r2*phi_.primitiveField()[f];
    forAll(r->boundaryFieldRef(), pi)
    {
        fvsPatchScalarField& pf = r->boundaryFieldRef()[pi];
        forAll(pf, i) pf[i] = d*alphaPhi_.boundaryField()[pi][i] + r2*phi_.boundaryField()[pi][i];
    }
    return tmp<surfaceScalarField>(r);
}

tmp<surfaceScalarField> solveVofEqu::getRhoPhi(const volScalarField& rho1, const volScalarField& rho2)
{
    // advection.H:346-361: fvc::interpolate(rho1 - rho2)*alphaPhi + fvc::interpolate(rho2)*phi
    volScalarField dRho("rho1-rho2", mesh_, 0.0);
    forAll(dRho.primitiveFieldRef(), i) dRho.primitiveFieldRef()[i] = rho1.primitiveField()[i] - rho2.primitiveField()[i];
    forAll(dRho.boundaryFieldRef(), pi)
    {
        forAll(dRho.boundaryFieldRef()[pi], i) dRho.boundaryFieldRef()[pi][i] = rho1.boundaryField()[pi][i] - rho2.boundaryField()[pi][i];
    }
    tmp<surfaceScalarField> td(fvc::interpolate(dRho)), t2(fvc::interpolate(rho2));
    surfaceScalarField* r = new surfaceScalarField("rhoPhi", mesh_, 0.0);
    forAll(r->primitiveFieldRef(), f)
        r->primitiveFieldRef()[f] = td().primitiveField()[f]*alphaPhi_.primitiveField()[f] + t2().primitiveField()[f]*phi_.primitiveField()[f];
    forAll(r->boundaryFieldRef(), pi)
    {
        fvsPatchScalarField& pf = r->boundaryFieldRef()[pi];
        forAll(pf, i) pf[i] = td().boundaryField()[pi][i]*alphaPhi_.boundaryField()[pi][i] + t2().boundaryField()[pi][i]*phi_.boundaryField()[pi][i];
    }
    return tmp<surfaceScalarField>(r);
}

scalar solveVofEqu::advectionTime() const
{
    double v = 0;
    svof_get_info(h_, SVOF_I_ADVECTION_TIME, &v);
    return v;
}

// * * * * * * * * * * * * * * * * reconstruction * * * * * * * * * * * * * * * //

reconstruction::reconstruction(volScalarField& alpha1, solveVofEqu& owner)
:
    IOdictionary(IOobject("reconstruction", alpha1.time().constant(), alpha1.db(), IOobject::NO_READ, IOobject::NO_WRITE)),
    mesh_(alpha1.mesh()),
    owner_(owner)
{}

void reconstruction::reconstruct()
{
    owner_.pushAlpha();          // the caller may have touched alpha1 since the last call (sub-cycling, PIMPLE)
    solveVofEqu::check(svof_reconstruct(owner_.h_), owner_.h_, "svof_reconstruct");
    double nMixed = 0;
    svof_get_info(owner_.h_, SVOF_I_N_MIXED, &nMixed);
    Info<< "SimPLIC::reconstruction: Number of mixed cells = "
        << returnReduce(label(nMixed), sumOp<label>()) << endl;                     // reconstruction.C:675
    if (owner_.mapAlphaFieldOn_)
    {
        owner_.interfaceNMapped_ = interfaceN();
        owner_.interfaceDMapped_ = interfaceD();
    }
}

plicSurface reconstruction::interface()
{
    plicSurface s;
    int64_t nP = 0, nF = 0;
    solveVofEqu::check(svof_plic_surface(owner_.h_, 0, 0, nullptr, nullptr, nullptr, &nP, &nF), owner_.h_, "svof_plic_surface");
    scalarField pts(3*label(nP));
    labelList off(label(nF) + 1, 0), cells(label(nF), 0);
    solveVofEqu::check(svof_plic_surface(owner_.h_, nP, nF, pts.data(), off.data(), cells.data(), &nP, &nF), owner_.h_, "svof_plic_surface");
    s.points.setSize(label(nP));
    forAll(s.points, i) s.points[i] = point(pts[3*i], pts[3*i + 1], pts[3*i + 2]);
    s.faces.setSize(label(nF));
    s.meshCells = cells;
    forAll(s.faces, i)
    {
        s.faces[i].setSize(off[i + 1] - off[i]);
        forAll(s.faces[i], k) s.faces[i][k] = off[i] + k;
    }
    return s;
}

plicSurface reconstruction::subCellFaces()
{
    plicSurface s;
    int64_t nP = 0, nF = 0, nFP = 0;
    solveVofEqu::check(svof_subcell_faces(owner_.h_, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, &nP, &nF, &nFP), owner_.h_,
                       "svof_subcell_faces");
    scalarField pts(3*label(nP));
    labelList off(label(nF) + 1, 0), fpts(label(nFP), 0), cells(label(nF), 0);
    solveVofEqu::check(svof_subcell_faces(owner_.h_, nP, nF, nFP, pts.data(), off.data(), fpts.data(), cells.data(), &nP, &nF, &nFP),
                       owner_.h_, "svof_subcell_faces");
    s.points.setSize(label(nP));
    forAll(s.points, i) s.points[i] = point(pts[3*i], pts[3*i + 1], pts[3*i + 2]);
    s.faces.setSize(label(nF));
    s.meshCells = cells;
    forAll(s.faces, i)
    {
        s.faces[i].setSize(off[i + 1] - off[i]);
        forAll(s.faces[i], k) s.faces[i][k] = fpts[off[i] + k];
    }
    return s;
}

static labelList labelField(svof_handle* h, int which, label cap)
{
    labelList l(cap, 0);
    const int64_t n = svof_get_field(h, which, l.data(), cap);
    l.setSize(label(n < 0 ? 0 : n));
    return l;
}
static scalarField realField(svof_handle* h, int which, label n)
{
    scalarField f(n, 0.0);
    svof_get_field(h, which, f.data(), n);
    return f;
}
static vectorField vecField(svof_handle* h, int which, label n)
{
    scalarField f(realField(h, which, 3*n));
    vectorField v(n);
    forAll(v, i) v[i] = vector(f[3*i], f[3*i + 1], f[3*i + 2]);
    return v;
}

labelList reconstruction::mixedCells() const { return labelField(owner_.h_, SVOF_F_MIXED_CELLS, owner_.nCellsDev_); }
labelList reconstruction::cellStatus() const { return labelField(owner_.h_, SVOF_F_CELL_STATUS, owner_.nCellsDev_); }
vectorField reconstruction::interfaceN() const { return vecField(owner_.h_, SVOF_F_INTERFACE_N, owner_.nCellsDev_); }
scalarField reconstruction::interfaceD() const { return realField(owner_.h_, SVOF_F_INTERFACE_D, owner_.nCellsDev_); }
vectorField reconstruction::interfaceC() const { return vecField(owner_.h_, SVOF_F_INTERFACE_C, owner_.nCellsDev_); }
vectorField reconstruction::interfaceS() const { return vecField(owner_.h_, SVOF_F_INTERFACE_S, owner_.nCellsDev_); }
scalarField reconstruction::faceFlatness() const { return realField(owner_.h_, SVOF_F_FACE_FLATNESS, owner_.nFacesDev_); }

scalar reconstruction::reconstructionTime() const
{
    double v = 0;
    svof_get_info(owner_.h_, SVOF_I_RECONSTRUCTION_TIME, &v);
    return v;
}

scalar reconstruction::alphaMappingTime() const
{
    double v = 0;
    svof_get_info(owner_.h_, SVOF_I_ALPHA_MAPPING_TIME, &v);
    return v;
}

} // End namespace SimPLIC
} // End namespace geometricVofExt
} // End namespace Foam
