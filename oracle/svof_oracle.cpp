// TEST INFRASTRUCTURE ONLY.
//
// CPU oracle of the SimPLIC volume-fraction transport step: a single-threaded
// restatement of the reference algorithm (src/SimPLIC/{cut,reconstruction,
// advection}) exported through the SAME C ABI as the product (include/svof.h),
// so parity tests drive both libraries with identical calls.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product never does.
//
// Build: oracle/build.py  ->  oracle/_build/libsvof_oracle.so
//        g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "ora_solver.hpp"

using namespace ora;

struct svof_handle {
    Solver s;
    std::string err;
    double lastReconMs = 0, lastAdvectMs = 0;
    double marks[8] = {0};
};

static thread_local std::string g_create_error = "";

static double nowSec()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

#define ORA_TRY(h) try {
#define ORA_CATCH(h, code)                       \
    }                                            \
    catch (const std::exception& e)              \
    {                                            \
        (h)->err = e.what();                     \
        return code;                             \
    }

extern "C" {

int svof_params_default(svof_params* p)
{
    if (!p) return SVOF_ERR_INVALID_ARG;
    std::memset(p, 0, sizeof(*p));
    p->mixed_cell_tol = 1e-8;   // reconstruction.C:502
    p->snap_tol = 0.0;          // advection.C:456
    p->iso_face_tol = 0.0;
    p->rdf_tol = 1e-6;          // reconstruction.C:513
    p->rdf_rel_tol = 0.1;       // :514
    p->n_alpha_bounds = 10;     // advection.C:455
    p->clip = 1;                // advection.C:457
    p->alpha_grad_scheme = 0;
    p->orientation_method = SVOF_ORIENT_ISO_ALPHA_GRAD;  // reconstruction.C:507
    p->split_warped_face = 0;   // :504
    p->map_alpha_field = 0;     // :516
    p->write_plic_fields = 0;   // :503
    p->rdf_iterations = 5;      // :515
    p->mixed_cell_tol_set = 0;
    return SVOF_OK;
}

static bool parseBool(const char* v, int32_t* out)
{
    // OpenFOAM Switch spellings
    static const char* T[] = {"true", "on", "yes", "y", "t", "1"};
    static const char* F[] = {"false", "off", "no", "n", "f", "0", "none"};
    for (const char* s : T)
        if (!std::strcmp(v, s)) { *out = 1; return true; }
    for (const char* s : F)
        if (!std::strcmp(v, s)) { *out = 0; return true; }
    return false;
}

int svof_params_set(svof_params* p, const char* key, const char* value)
{
    if (!p || !key || !value) return SVOF_ERR_INVALID_ARG;
    std::string k(key), v(value);
    while (!v.empty() && (v.back() == ';' || v.back() == ' ')) v.pop_back();
    char* end = nullptr;
    auto num = [&](double* out) {
        *out = std::strtod(v.c_str(), &end);
        return end != v.c_str() && *end == '\0';
    };
    double d;
    int32_t b;
    if (k == "mixedCellTol") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; p->mixed_cell_tol = d; p->mixed_cell_tol_set = 1; return SVOF_OK; }
    if (k == "surfCellTol") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; if (!p->mixed_cell_tol_set) p->mixed_cell_tol = d; return SVOF_OK; }
    if (k == "isoFaceTol") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; p->iso_face_tol = d; return SVOF_OK; }
    if (k == "snapTol") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; p->snap_tol = d; return SVOF_OK; }
    if (k == "tol") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; p->rdf_tol = d; return SVOF_OK; }
    if (k == "relTol") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; p->rdf_rel_tol = d; return SVOF_OK; }
    if (k == "nAlphaBounds") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; p->n_alpha_bounds = int32_t(d); return SVOF_OK; }
    if (k == "iterations") { if (!num(&d)) return SVOF_ERR_INVALID_ARG; p->rdf_iterations = int32_t(d); return SVOF_OK; }
    if (k == "clip") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->clip = b; return SVOF_OK; }
    if (k == "splitWarpedFace") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->split_warped_face = b; return SVOF_OK; }
    if (k == "mapAlphaField") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->map_alpha_field = b; return SVOF_OK; }
    if (k == "writePlicFields") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->write_plic_fields = b; return SVOF_OK; }
    if (k == "orientationMethod") {
        // reconstruction.C:60-68; anything else is the FatalError of :610-625
        if (v == "alphaGrad") p->orientation_method = SVOF_ORIENT_ALPHA_GRAD;
        else if (v == "isoAlphaGrad" || v == "LS") p->orientation_method = SVOF_ORIENT_ISO_ALPHA_GRAD;
        else if (v == "isoRDF" || v == "RDF") p->orientation_method = SVOF_ORIENT_ISO_RDF;
        else return SVOF_ERR_BAD_CONFIG;
        return SVOF_OK;
    }
    if (k == "gradSchemes" || k == "grad(alpha1)" || k == "gradScheme") {
        // the caller's fvSchemes entry for grad(alpha1), read by fvc::grad(alpha1_, "grad(alpha1)") (reconstruction.C:78)
        if (v == "Gauss linear" || v == "linear") p->alpha_grad_scheme = 0;
        else if (v == "Gauss pointLinear" || v == "pointLinear") p->alpha_grad_scheme = 1;
        else return SVOF_ERR_BAD_CONFIG;
        return SVOF_OK;
    }
    // caller-side keys living in the same dictionary
    if (k == "nAlphaSubCycles" || k == "cAlpha" || k == "period" || k == "reverseTime" || k == "nAlphaCorr") return SVOF_OK;
    return SVOF_ERR_INVALID_ARG;
}

int svof_create(const svof_mesh* mesh, const svof_params* params, const svof_comm* comm, svof_handle** out)
{
    if (!mesh || !params || !out) { g_create_error = "svof_create: null argument"; return SVOF_ERR_INVALID_ARG; }
    if (comm && comm->world_size > 1) { g_create_error = "the CPU oracle is single-domain"; return SVOF_ERR_UNSUPPORTED; }
    svof_handle* h = new (std::nothrow) svof_handle;
    if (!h) return SVOF_ERR_INVALID_ARG;
    try {
        h->s.init(*mesh, *params);
    } catch (const std::exception& e) {
        g_create_error = e.what();
        delete h;
        return SVOF_ERR_BAD_MESH;
    }
    *out = h;
    return SVOF_OK;
}

int svof_destroy(svof_handle* h)
{
    delete h;
    return SVOF_OK;
}

const char* svof_last_error(const svof_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int svof_set_alpha(svof_handle* h, const double* alpha)
{
    if (!h || !alpha) return SVOF_ERR_INVALID_ARG;
    h->s.alpha.assign(alpha, alpha + h->s.mesh.nCells);
    h->s.alphaOld = h->s.alpha;
    h->s.correctAlphaBCs();
    h->s.haveAlpha = true;
    return SVOF_OK;
}

int svof_set_phi(svof_handle* h, const double* phi)
{
    if (!h || !phi) return SVOF_ERR_INVALID_ARG;
    h->s.phi.assign(phi, phi + h->s.mesh.nFaces);
    h->s.havePhi = true;
    return SVOF_OK;
}

int svof_set_U(svof_handle* h, const double* U, const double* Ub)
{
    if (!h || !U) return SVOF_ERR_INVALID_ARG;
    Solver& s = h->s;
    for (label c = 0; c < s.mesh.nCells; ++c) s.U[c] = vec(U[3 * c], U[3 * c + 1], U[3 * c + 2]);
    const label nBF = s.mesh.nBoundaryFaces();
    for (label b = 0; b < nBF; ++b) s.Ub[b] = Ub ? vec(Ub[3 * b], Ub[3 * b + 1], Ub[3 * b + 2]) : vec();
    s.haveU = true;
    return SVOF_OK;
}

int svof_reconstruct(svof_handle* h)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (!h->s.haveAlpha) { h->err = "svof_reconstruct: alpha not set"; return SVOF_ERR_STATE; }
    const double t0 = nowSec();
    ORA_TRY(h)
    h->s.reconstruct();
    ORA_CATCH(h, SVOF_ERR_INVALID_ARG)
    const double dt = nowSec() - t0;
    h->s.reconstructionTime += dt;
    h->lastReconMs = dt * 1e3;
    return SVOF_OK;
}

int svof_advect(svof_handle* h, double dt, const double* Sp, const double* Su)
{
    if (!h || !(dt > 0)) return SVOF_ERR_INVALID_ARG;
    if (!h->s.haveAlpha || !h->s.havePhi || !h->s.haveU) { h->err = "svof_advect: alpha/phi/U not set"; return SVOF_ERR_STATE; }
    const double t0 = nowSec();
    ORA_TRY(h)
    h->s.advect(dt, Sp, Su);
    ORA_CATCH(h, SVOF_ERR_INVALID_ARG)
    const double el = nowSec() - t0;
    h->s.advectionTime += el;
    h->lastAdvectMs = el * 1e3;
    return SVOF_OK;
}

int svof_step_host(svof_handle* h, double dt, const double* phi, const double* U, const double* Ub, double* alpha_out,
                   double* alpha_phi_out)
{
    int rc;
    if ((rc = svof_set_phi(h, phi))) return rc;
    if ((rc = svof_set_U(h, U, Ub))) return rc;
    if ((rc = svof_reconstruct(h))) return rc;
    if ((rc = svof_advect(h, dt, nullptr, nullptr))) return rc;
    if (alpha_out) std::memcpy(alpha_out, h->s.alpha.data(), sizeof(double) * h->s.mesh.nCells);
    if (alpha_phi_out) std::memcpy(alpha_phi_out, h->s.alphaPhi.data(), sizeof(double) * h->s.mesh.nFaces);
    return SVOF_OK;
}

static int64_t copyOut(const void* src, int64_t n, size_t esz, void* dst, int64_t cap)
{
    if (cap < n) return SVOF_ERR_INVALID_ARG;
    std::memcpy(dst, src, size_t(n) * esz);
    return n;
}

int64_t svof_get_field(svof_handle* h, int which, void* dst, int64_t capacity)
{
    if (!h || !dst) return SVOF_ERR_INVALID_ARG;
    Solver& s = h->s;
    const int64_t nC = s.mesh.nCells, nF = s.mesh.nFaces, nM = int64_t(s.mixedCells.size());
    switch (which) {
        case SVOF_F_ALPHA: return copyOut(s.alpha.data(), nC, 8, dst, capacity);
        case SVOF_F_ALPHA_PHI: return copyOut(s.alphaPhi.data(), nF, 8, dst, capacity);
        case SVOF_F_DVF: return copyOut(s.dVf.data(), nF, 8, dst, capacity);
        case SVOF_F_INTERFACE_N: return copyOut(s.interfaceN.data(), 3 * nC, 8, dst, capacity);
        case SVOF_F_INTERFACE_D: return copyOut(s.interfaceD.data(), nC, 8, dst, capacity);
        case SVOF_F_INTERFACE_C: return copyOut(s.interfaceC.data(), 3 * nC, 8, dst, capacity);
        case SVOF_F_INTERFACE_S: return copyOut(s.interfaceS.data(), 3 * nC, 8, dst, capacity);
        case SVOF_F_MIXED_CELLS: return copyOut(s.mixedCells.data(), nM, 4, dst, capacity);
        case SVOF_F_CELL_STATUS: return copyOut(s.cellStatus.data(), nM, 4, dst, capacity);
        case SVOF_F_FACE_FLATNESS: return copyOut(s.mesh.faceFlatness.data(), nF, 8, dst, capacity);
        case SVOF_F_CF: return copyOut(s.mesh.Cf.data(), 3 * nF, 8, dst, capacity);
        case SVOF_F_SF: return copyOut(s.mesh.Sf.data(), 3 * nF, 8, dst, capacity);
        case SVOF_F_C: return copyOut(s.mesh.C.data(), 3 * nC, 8, dst, capacity);
        case SVOF_F_V: return copyOut(s.mesh.V.data(), nC, 8, dst, capacity);
        case SVOF_F_ALPHA_BOUNDARY: return copyOut(s.alphaB.data(), nF - s.mesh.nInternalFaces, 8, dst, capacity);
        case SVOF_F_UN0: return copyOut(s.Un0.data(), int64_t(s.Un0.size()), 8, dst, capacity);
    }
    return SVOF_ERR_INVALID_ARG;
}

int svof_get_info(svof_handle* h, int which, double* out)
{
    if (!h || !out) return SVOF_ERR_INVALID_ARG;
    Solver& s = h->s;
    switch (which) {
        case SVOF_I_N_MIXED: *out = double(s.mixedCells.size()); return SVOF_OK;
        case SVOF_I_MIN_ALPHA_BEFORE: *out = s.minBefore; return SVOF_OK;
        case SVOF_I_MAX_ALPHA_M1_BEFORE: *out = s.maxM1Before; return SVOF_OK;
        case SVOF_I_MIN_ALPHA_AFTER: *out = s.minAfter; return SVOF_OK;
        case SVOF_I_MAX_ALPHA_M1_AFTER: *out = s.maxM1After; return SVOF_OK;
        case SVOF_I_N_BOUND_SWEEPS: *out = s.nSweeps; return SVOF_OK;
        case SVOF_I_RECONSTRUCTION_TIME: *out = s.reconstructionTime; return SVOF_OK;
        case SVOF_I_ADVECTION_TIME: *out = s.advectionTime; return SVOF_OK;
        case SVOF_I_ALPHA_MAPPING_TIME: *out = h->s.alphaMappingTime; return SVOF_OK;
        case SVOF_I_VOLUME: *out = s.volume(); return SVOF_OK;
        case SVOF_I_GPU_LAUNCHES: *out = 0; return SVOF_OK;
        case SVOF_I_FLATNESS_MIN: *out = s.mesh.flatMin; return SVOF_OK;
        case SVOF_I_FLATNESS_MAX: *out = s.mesh.flatMax; return SVOF_OK;
        case SVOF_I_FLATNESS_AVG: *out = s.mesh.flatAvg; return SVOF_OK;
        case SVOF_I_DEVICE_BYTES: *out = 0; return SVOF_OK;
        case SVOF_I_ERROR_FLAGS: *out = 0; return SVOF_OK;
        case SVOF_I_DENSE_KERNEL_MS: *out = 0; return SVOF_OK;
        case SVOF_I_DENSE_KERNEL_LAUNCHES: *out = 0; return SVOF_OK;
        case SVOF_I_N_NEAR: *out = 0; return SVOF_OK;
        case SVOF_I_H2D_BYTES: *out = 0; return SVOF_OK;
        case SVOF_I_D2H_BYTES: *out = 0; return SVOF_OK;
        case SVOF_I_RDF_ITERATIONS: *out = double(s.isoRDFIterationsDone); return SVOF_OK;
    }
    return SVOF_ERR_INVALID_ARG;
}

int svof_device_ptr(svof_handle*, int, void**) { return SVOF_ERR_UNSUPPORTED; }
int svof_device_touch(svof_handle*, int) { return SVOF_ERR_UNSUPPORTED; }
int svof_scatter_alpha_device(svof_handle*, const int32_t*, const double*, int64_t) { return SVOF_ERR_UNSUPPORTED; }
int svof_set_phi_device(svof_handle*, const void*) { return SVOF_ERR_UNSUPPORTED; }
int svof_set_U_device(svof_handle*, const void*, const void*) { return SVOF_ERR_UNSUPPORTED; }
int svof_set_option(svof_handle* h, const char* name, int) { return (h && name) ? SVOF_OK : SVOF_ERR_INVALID_ARG; }
int svof_get_stream(svof_handle*, void**) { return SVOF_ERR_UNSUPPORTED; }
int svof_synchronize(svof_handle*) { return SVOF_OK; }
int svof_mark(svof_handle* h, int slot)
{
    if (!h || slot < 0 || slot > 7) return SVOF_ERR_INVALID_ARG;
    h->marks[slot] = nowSec();
    return SVOF_OK;
}
int svof_elapsed_ms(svof_handle* h, int a, int b, double* ms)
{
    if (!h || !ms || a < 0 || a > 7 || b < 0 || b > 7) return SVOF_ERR_INVALID_ARG;
    *ms = (h->marks[b] - h->marks[a]) * 1e3;
    return SVOF_OK;
}
int svof_last_step_ms(svof_handle* h, double* r, double* a)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (r) *r = h->lastReconMs;
    if (a) *a = h->lastAdvectMs;
    return SVOF_OK;
}

int svof_host_alloc(int64_t bytes, void** out)
{
    if (!out || bytes <= 0) return SVOF_ERR_INVALID_ARG;
    *out = std::malloc(size_t(bytes));
    return *out ? SVOF_OK : SVOF_ERR_INVALID_ARG;
}
int svof_host_free(void* p)
{
    std::free(p);
    return SVOF_OK;
}

// ---- geometry primitives ------------------------------------------------------

int svof_cut_faces(svof_handle* h, int32_t n_polys, int32_t n_verts, const double* pts, const double* normals,
                   const double* dists, int32_t* status, double* centres, double* areas)
{
    if (!h || n_verts < 3) return SVOF_ERR_INVALID_ARG;
    cutFace cf(h->s.mesh);
    std::vector<point> fPts(n_verts);
    for (int32_t i = 0; i < n_polys; ++i) {
        for (int32_t k = 0; k < n_verts; ++k) {
            const double* q = pts + (size_t(i) * n_verts + k) * 3;
            fPts[k] = point(q[0], q[1], q[2]);
        }
        status[i] = cf.calcSubFace(fPts, vec(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]), dists[i]);
        centres[3 * i] = cf.subFaceCentre().x; centres[3 * i + 1] = cf.subFaceCentre().y; centres[3 * i + 2] = cf.subFaceCentre().z;
        areas[3 * i] = cf.subFaceArea().x; areas[3 * i + 1] = cf.subFaceArea().y; areas[3 * i + 2] = cf.subFaceArea().z;
    }
    return SVOF_OK;
}

int svof_cut_cells(svof_handle* h, int32_t n, const int32_t* cells, const double* normals, const double* dists,
                   int32_t* status, double* vof, double* sub_volume, double* ic, double* ia)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    cutCell cc(h->s.mesh);
    for (int32_t i = 0; i < n; ++i) {
        if (cells[i] < 0 || cells[i] >= h->s.mesh.nCells) return SVOF_ERR_INVALID_ARG;
        status[i] = cc.calcSubCell(cells[i], vec(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]), dists[i], false);
        vof[i] = cc.volumeOfFluid();
        sub_volume[i] = cc.subCellVolume();
        ic[3 * i] = cc.interfaceCentre().x; ic[3 * i + 1] = cc.interfaceCentre().y; ic[3 * i + 2] = cc.interfaceCentre().z;
        ia[3 * i] = cc.interfaceArea().x; ia[3 * i + 1] = cc.interfaceArea().y; ia[3 * i + 2] = cc.interfaceArea().z;
    }
    return SVOF_OK;
}

int svof_step_device(svof_handle* h, double dt)
{
    const int rc = svof_reconstruct(h);
    return rc ? rc : svof_advect(h, dt, nullptr, nullptr);
}

int svof_plic_surface(svof_handle* h, int64_t cap_points, int64_t cap_faces, double* points, int32_t* face_offsets, int32_t* cells,
                      int64_t* n_points, int64_t* n_faces)
{
    if (!h || !n_points || !n_faces) return SVOF_ERR_INVALID_ARG;
    Solver& s = h->s;
    cutCell cc(s.mesh);
    std::vector<point> poly, pts;
    std::vector<int32_t> off(1, 0), cl;
    for (size_t i = 0; i < s.mixedCells.size(); ++i) {
        const label c = s.mixedCells[i];
        if (cc.calcSubCell(c, s.interfaceN[c], s.interfaceD[c], false) != 0) continue;  // reconstruction.C:800-806
        cc.interfacePolygon(poly);
        if (poly.empty()) continue;
        pts.insert(pts.end(), poly.begin(), poly.end());
        off.push_back(int32_t(pts.size()));
        cl.push_back(c);
    }
    *n_points = int64_t(pts.size());
    *n_faces = int64_t(cl.size());
    if (!points) return SVOF_OK;
    if (cap_points < *n_points || cap_faces < *n_faces || !face_offsets || !cells) return SVOF_ERR_CAPACITY;
    for (size_t i = 0; i < pts.size(); ++i) { points[3 * i] = pts[i].x; points[3 * i + 1] = pts[i].y; points[3 * i + 2] = pts[i].z; }
    std::copy(off.begin(), off.end(), face_offsets);
    std::copy(cl.begin(), cl.end(), cells);
    return SVOF_OK;
}

// reconstruction::subCellFaces() (reconstruction.C:838-891)
int svof_subcell_faces(svof_handle* h, int64_t cap_points, int64_t cap_faces, int64_t cap_face_points, double* points,
                       int32_t* face_offsets, int32_t* face_points, int32_t* face_cell, int64_t* n_points, int64_t* n_faces,
                       int64_t* n_face_points)
{
    if (!h || !n_points || !n_faces || !n_face_points) return SVOF_ERR_INVALID_ARG;
    Solver& s = h->s;
    cutCell cc(s.mesh);
    cc.collectSubCell(true);
    std::vector<point> pts, cp;
    std::vector<std::vector<label>> cf;
    std::vector<int32_t> off(1, 0), fpts, fcell;
    for (size_t i = 0; i < s.mixedCells.size(); ++i) {
        const label c = s.mixedCells[i];
        if (cc.calcSubCell(c, s.interfaceN[c], s.interfaceD[c], false) != 0) continue;
        cc.subCellPointsAndFaces(cp, cf);
        const int32_t base = int32_t(pts.size());
        for (const std::vector<label>& f : cf) {
            for (label v : f) fpts.push_back(base + v);
            off.push_back(int32_t(fpts.size()));
            fcell.push_back(c);
        }
        pts.insert(pts.end(), cp.begin(), cp.end());
    }
    *n_points = int64_t(pts.size());
    *n_faces = int64_t(fcell.size());
    *n_face_points = int64_t(fpts.size());
    if (!points) return SVOF_OK;
    if (cap_points < *n_points || cap_faces < *n_faces || cap_face_points < *n_face_points || !face_offsets || !face_points || !face_cell)
        return SVOF_ERR_CAPACITY;
    for (size_t i = 0; i < pts.size(); ++i) { points[3 * i] = pts[i].x; points[3 * i + 1] = pts[i].y; points[3 * i + 2] = pts[i].z; }
    std::copy(off.begin(), off.end(), face_offsets);
    std::copy(fpts.begin(), fpts.end(), face_points);
    std::copy(fcell.begin(), fcell.end(), face_cell);
    return SVOF_OK;
}

int svof_find_signed_distance(svof_handle* h, int32_t n, const int32_t* cells, const double* alphas, const double* normals,
                              int32_t* status, double* dists, double* ic, double* ia)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    cutCell cc(h->s.mesh);
    for (int32_t i = 0; i < n; ++i) {
        if (cells[i] < 0 || cells[i] >= h->s.mesh.nCells) return SVOF_ERR_INVALID_ARG;
        scalar D = 0;
        vec C, S;
        status[i] = cc.findSignedDistance(cells[i], alphas[i], vec(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]),
                                          h->s.prm.split_warped_face != 0, D, C, S);
        dists[i] = D;
        ic[3 * i] = C.x; ic[3 * i + 1] = C.y; ic[3 * i + 2] = C.z;
        ia[3 * i] = S.x; ia[3 * i + 1] = S.y; ia[3 * i + 2] = S.z;
    }
    return SVOF_OK;
}

int svof_face_fluxes(svof_handle* h, int32_t n, const int32_t* faces, const double* normals, const double* dists,
                     const double* Un0, double dt, const double* phi, double* dVf)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    cutFace cf(h->s.mesh);
    for (int32_t i = 0; i < n; ++i) {
        const label f = faces[i];
        if (f < 0 || f >= h->s.mesh.nFaces) return SVOF_ERR_INVALID_ARG;
        dVf[i] = cf.timeIntegratedFaceFlux(f, vec(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]), dists[i], Un0[i], dt,
                                           phi[i], h->s.mesh.magSf[f]);
    }
    return SVOF_OK;
}

// ---- changing meshes ------------------------------------------------------------------------------------
int svof_update_points(svof_handle* h, const double* points, const double* Cf, const double* Sf, const double* C, const double* V)
{
    if (!h || !points) return SVOF_ERR_INVALID_ARG;
    if ((Cf || Sf || C || V) && !(Cf && Sf && C && V)) return SVOF_ERR_INVALID_ARG;
    ORA_TRY(h)
    h->s.mesh.movePoints(points, Cf, Sf, C, V);
    return SVOF_OK;
    ORA_CATCH(h, SVOF_ERR_BAD_MESH)
}

int svof_update_mesh(svof_handle* h, const svof_mesh* mesh)
{
    if (!h || !mesh) return SVOF_ERR_INVALID_ARG;
    ORA_TRY(h)
    const svof_params p = h->s.prm;
    h->s.~Solver();
    new (&h->s) Solver();
    h->s.init(*mesh, p);
    return SVOF_OK;
    ORA_CATCH(h, SVOF_ERR_BAD_MESH)
}

int svof_set_interface(svof_handle* h, const double* N, const double* D)
{
    if (!h || !N || !D) return SVOF_ERR_INVALID_ARG;
    for (label c = 0; c < h->s.mesh.nCells; ++c) {
        h->s.interfaceN[c] = vec(N[3 * c], N[3 * c + 1], N[3 * c + 2]);
        h->s.interfaceD[c] = D[c];
    }
    return SVOF_OK;
}

int svof_set_cell_types(svof_handle* h, const int32_t* cell_types)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (!cell_types) h->s.cellTypes.clear();
    else h->s.cellTypes.assign(cell_types, cell_types + h->s.mesh.nCells);
    return SVOF_OK;
}

int svof_map_alpha_field(svof_handle* h, double lower, double upper)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (!h->s.haveAlpha) { h->err = "svof_map_alpha_field: alpha not set"; return SVOF_ERR_STATE; }
    ORA_TRY(h)
    const double t0 = nowSec();
    h->s.mapAlphaField(lower, upper);
    h->s.alphaMappingTime += nowSec() - t0;
    return SVOF_OK;
    ORA_CATCH(h, SVOF_ERR_INVALID_ARG)
}

// decomposed runs: the CPU oracle is single-domain.  Sub-domain extraction and partitioning are host utilities of the
// product library (no CUDA involved); the tests drive per-rank oracles on the sub-meshes it produces and exchange the
// ghost values themselves (gloo).
int svof_partition_rcb(const svof_mesh*, const double*, int32_t, int32_t*) { return SVOF_ERR_UNSUPPORTED; }
int svof_decompose(const svof_mesh*, const int32_t*, int32_t, int32_t, svof_submesh**) { return SVOF_ERR_UNSUPPORTED; }
int svof_submesh_mesh(const svof_submesh*, svof_mesh*) { return SVOF_ERR_UNSUPPORTED; }
int svof_submesh_maps(const svof_submesh*, int32_t*, const int32_t**, const int32_t**, const int32_t**, const int32_t**,
                      const int32_t**, const int32_t**) { return SVOF_ERR_UNSUPPORTED; }
int svof_submesh_free(svof_submesh*) { return SVOF_ERR_UNSUPPORTED; }
const char* svof_decomp_last_error(void) { return "the CPU oracle has no decomposition"; }
int svof_comm_unique_id(void*) { return SVOF_ERR_UNSUPPORTED; }
int svof_halo_setup(svof_handle*, const int32_t*, const int32_t*) { return SVOF_ERR_UNSUPPORTED; }
int svof_halo_exchange(svof_handle*) { return SVOF_ERR_UNSUPPORTED; }
int svof_halo_setup_faces(svof_handle*, const int32_t*, const int32_t*, const int32_t*) { return SVOF_ERR_UNSUPPORTED; }
int svof_halo_exchange_inputs(svof_handle*) { return SVOF_ERR_UNSUPPORTED; }
int svof_submesh_face_maps(const svof_submesh*, const int32_t**, const int32_t**) { return SVOF_ERR_UNSUPPORTED; }

}  // extern "C"
