// svof_geom.cuh -- device geometry of the SimPLIC step (sm_100a, FP64).
//
// What the reference computes with two stateful scratch classes and dozens of
// heap allocations per cut cell (cutFace.C / cutCell.C) is evaluated here by one
// thread per polyhedron entirely in registers + thread-local scratch:
//   clipFace            <- cutFace::calcSubFace + calcSubFaceCentreAndArea (cutFace.C:37-96,136-259)
//   timeIntegratedArea  <- cutFace::timeIntegratedArea                     (cutFace.C:392-506)
//   faceFlux            <- cutFace::timeIntegratedFaceFlux                 (cutFace.C:262-389)
//   subCell             <- cutCell::calcSubCell + calcInterfaceCentreAndArea
//                          + calcSubCellCentreAndVolume (cutCell.C:37-137,343-542), including the
//                          splitWarpedFace local triangulation (:140-236) generated on the fly
//   signedDistance      <- cutCell::findSignedDistance                     (cutCell.C:611-799)
// No sub-cell point/face lists are built (the reference appends them at
// cutCell.C:382-393,443-454 and never reads them on this path).
// Arithmetic follows the reference expression by expression (see svof_math.cuh).
#pragma once
#include "svof_math.cuh"

namespace svof {

enum {
    SVERR_FACE_VERTS = 1,   // a face has more vertices than the compiled cap
    SVERR_CELL_FACES = 2,   // a cell has more (local) faces than the cap
    SVERR_CELL_POINTS = 4,  // a cell has more (local) points than the cap
    SVERR_IFACE_POINTS = 8, // a clipped face produced more interface points than the cap
    SVERR_STENCIL = 16,     // LS stencil larger than the cap
    SVERR_LIST = 32         // a device work list overflowed its allocation
};

// Device view of the mesh: CSR/SoA, int32 labels, f64 scalars (DESIGN.md "data layout").
struct MeshDev {
    int nPoints, nFaces, nIF, nCells, nBF;
    int maxFV, maxLocalFaces, maxLocalPts;   // mesh maxima: vertices per (local) face, local faces / points per cell (with splitWarpedFace: triangulated)
    const double* points;       // [3*nPoints]
    const int* faceOff;         // [nFaces+1]
    const int* facePts;
    const int* owner;           // [nFaces]
    const int* neighbour;       // [nIF]
    const double* Cf;           // [3*nFaces]
    const double* Sf;           // [3*nFaces]
    const double* magSf;        // [nFaces]
    const double* C;            // [3*nCells]
    const double* V;            // [nCells]
    const double* flat;         // [nFaces] face flatness
    const int* cellOff;         // [nCells+1]
    const int* cellFaces;       // primitiveMesh::cells() order (owned asc, then neighbour-side asc)
    const int2* cellAsc;        // same rows in ASCENDING face order: {face | flip<<31, other cell or -1-bFace}
    const int* cellPtOff;       // [nCells+1]
    const int* cellPts;         // ascending point labels
    const int* ptCellOff;       // [nPoints+1]
    const int* ptCells;         // ascending cell labels
    const int* ptBFOff;         // [nPoints+1]
    const int* ptBFaces;        // ascending boundary-face index
    const unsigned char* bKind; // [nBF] svof_patch_kind
    const unsigned char* isPatchPoint;  // [nPoints]
    const unsigned char* tetBase;       // [nFaces] polyMesh::tetBasePtIs
};

template <int MAXFV_, int MAXCF_, int MAXCP_>
struct Caps {
    static constexpr int MAXFV = MAXFV_;  // vertices per face
    static constexpr int MAXCF = MAXCF_;  // (local) faces per cell
    static constexpr int MAXCP = MAXCP_;  // (local) points per cell
    static constexpr int MAXIP = 4;       // interface points per clipped face
    static constexpr int MAXEP = 2 * MAXCF_ + 8;
};

// ---- face::centre / face::areaNormal (OF, recalled) on a local polygon ------------
__device__ __forceinline__ d3 faceCentreOF(const d3* p, int n)
{
    if (n == 3) return (1.0 / 3.0) * (p[0] + p[1] + p[2]);
    d3 cp = zero3();
    for (int i = 0; i < n; ++i) cp += p[i];
    cp /= double(n);
    double sumA = 0;
    d3 sumAc = zero3();
    for (int i = 0; i < n; ++i) {
        const d3 nx = p[(i + 1 == n) ? 0 : i + 1];
        const d3 ttc = p[i] + nx + cp;
        const double ta = mag(cross(p[i] - cp, nx - cp));
        sumA += ta;
        sumAc += ta * ttc;
    }
    if (sumA > SV_VSMALL) return sumAc / (3.0 * sumA);
    return cp;
}
__device__ __forceinline__ d3 faceAreaNormalOF(const d3* p, int n)
{
    if (n == 3) return 0.5 * cross(p[1] - p[0], p[2] - p[0]);
    d3 cp = zero3();
    for (int i = 0; i < n; ++i) cp += p[i];
    cp /= double(n);
    d3 a = zero3();
    for (int i = 0; i < n; ++i) {
        const d3 nx = (i < n - 1) ? p[i + 1] : p[0];
        a += 0.5 * cross(nx - p[i], cp - p[i]);
    }
    return a;
}

// cutFace::calcSubFaceCentreAndArea (cutFace.C:37-96)
__device__ __forceinline__ void subFaceCentreAndArea(const d3* sp, int np, d3& centre, d3& area)
{
    if (np == 3) {
        centre = (1.0 / 3.0) * (sp[0] + sp[1] + sp[2]);
        area = 0.5 * cross(sp[1] - sp[0], sp[2] - sp[0]);
        return;
    }
    d3 sumN = zero3(), sumAc = zero3();
    double sumA = 0.0;
    d3 fC = sp[0];
    for (int i = 1; i < np; ++i) fC += sp[i];
    fC /= double(np);
    for (int i = 0; i < np; ++i) {
        const d3 nx = sp[(i + 1 == np) ? 0 : i + 1];
        const d3 c = sp[i] + nx + fC;
        const d3 n = cross(nx - sp[i], fC - sp[i]);
        const double a = mag(n);
        sumN += n;
        sumA += a;
        sumAc += a * c;
    }
    if (sumA < SV_ROOTVSMALL) {
        centre = fC;
        area = zero3();
    } else {
        centre = (1.0 / 3.0) * sumAc / sumA;
        area = 0.5 * sumN;
    }
}

// cutFace::calcSubFace (cutFace.C:136-259).  ip/nip: interface points (valid for status 0).
template <class CP>
__device__ __noinline__ int clipFace(const d3* fp, int nv, const d3& n, double D, d3& centre, d3& area, d3* ip, int& nip, int& err,
                                     d3* spOut = nullptr, int* npOut = nullptr)
{
    double s[CP::MAXFV];
    int nSub = 0, first = -1;
    for (int i = 0; i < nv; ++i) {
        double si = dot(fp[i], n) + D;
        if (fabs(si) < SV_TSMALL) si += sgn(si) * SV_TSMALL;
        s[i] = si;
        if (si < 0.0) {
            nSub++;
            if (first < 0) first = i;
        }
    }
    nip = 0;
    if (nSub == nv) {
        centre = faceCentreOF(fp, nv);
        area = faceAreaNormalOF(fp, nv);
        return -1;
    }
    if (nSub == 0) {
        centre = zero3();
        area = zero3();
        return 1;
    }
    d3 sp[2 * CP::MAXFV];
    int np = 0;
    int cur = first;
    for (int i = 0; i < nv; ++i) {
        const int nxt = (cur + 1 == nv) ? 0 : cur + 1;
        if (s[cur] < 0) sp[np++] = fp[cur];
        if ((s[cur] * s[nxt]) < 0) {
            const double w = s[cur] / (s[cur] - s[nxt]);
            const d3 cp = fp[cur] + w * (fp[nxt] - fp[cur]);
            sp[np++] = cp;
            if (nip < CP::MAXIP) ip[nip] = cp; else err |= SVERR_IFACE_POINTS;
            nip++;
        }
        cur = nxt;
    }
    if (nip > CP::MAXIP) nip = CP::MAXIP;
    if (np >= 3) {
        subFaceCentreAndArea(sp, np, centre, area);
        if (spOut) {  // cutFace::subFacePoints(), for reconstruction::subCellFaces()
            for (int q = 0; q < np; ++q) spOut[q] = sp[q];
            *npOut = np;
        }
        return 0;
    }
    centre = faceCentreOF(fp, nv);
    area = faceAreaNormalOF(fp, nv);
    return -1;
}

template <class CP>
__device__ __forceinline__ int loadFace(const MeshDev& m, int f, d3* fp, int& err)
{
    const int o = __ldg(m.faceOff + f);
    int nv = __ldg(m.faceOff + f + 1) - o;
    if (nv > CP::MAXFV) {
        err |= SVERR_FACE_VERTS;
        nv = CP::MAXFV;
    }
    for (int k = 0; k < nv; ++k) fp[k] = ld3(m.points, __ldg(m.facePts + o + k));
    return nv;
}

// small stable insertion sorts on VALUES (the reference sorts indices with a stable sort and only
// ever reads the values back through them, so ties are indistinguishable)
__device__ __forceinline__ void sortAscending(double* v, int n)
{
    for (int i = 1; i < n; ++i) {
        const double x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = x;
    }
}
__device__ __forceinline__ void sortDescending(double* v, int n)
{
    for (int i = 1; i < n; ++i) {
        const double x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] < x) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = x;
    }
}

// cutFace::timeIntegratedArea (cutFace.C:392-506)
template <class CP>
__device__ __noinline__ double timeIntegratedArea(const d3* fp, int nv, const d3& n, double D, const double* pTimes, double Un0,
                                     double dt, double magSf, int& err)
{
    double st[CP::MAXFV];
    for (int i = 0; i < nv; ++i) st[i] = pTimes[i];
    sortAscending(st, nv);
    const double firstTime = st[0], lastTime = st[nv - 1];
    if (lastTime <= 0.0) return magSf * dt * pos0(Un0);
    if (firstTime >= dt) return magSf * dt * (1.0 - pos0(Un0));

    double tIntArea = 0.0;
    double times[CP::MAXFV + 2];
    int nt = 0;
    double prevTime = 0.0;
    double subAreaOld = 0.0, subAreaNew = 0.0, subAreaMid = 0.0;
    d3 c, a, ipd[CP::MAXIP];
    int nip;
    if (firstTime > 0.0) {
        subAreaOld = magSf * (1.0 - pos0(Un0));
        tIntArea = subAreaOld * firstTime;
        times[nt++] = firstTime;
        prevTime = firstTime;
    } else {
        times[nt++] = 0.0;
        prevTime = 0.0;
        clipFace<CP>(fp, nv, n, D, c, a, ipd, nip, err);
        subAreaOld = mag(a);
    }
    const double smallTime = dmax(SV_TSMALL / fabs(Un0), SV_TSMALL);
    for (int ti = 0; ti < nv; ++ti) {
        const double timeI = st[ti];
        if (timeI > (prevTime + smallTime) && timeI < dt) {
            times[nt++] = timeI;
            prevTime = timeI;
        }
    }
    if (lastTime > dt) {
        times[nt++] = dt;
    } else {
        tIntArea += magSf * (dt - lastTime) * pos0(Un0);
    }
    for (int k = 0; k < nt - 1; ++k) {
        const double tauOld = times[k], tauNew = times[k + 1];
        const double deltaTau = 0.5 * (tauNew - tauOld);
        clipFace<CP>(fp, nv, n, D - tauNew * Un0, c, a, ipd, nip, err);
        subAreaNew = mag(a);
        clipFace<CP>(fp, nv, n, D - (tauOld + deltaTau) * Un0, c, a, ipd, nip, err);
        subAreaMid = mag(a);
        tIntArea += (deltaTau / 3.0) * (subAreaOld + 4.0 * subAreaMid + subAreaNew);  // Simpson
        subAreaOld = subAreaNew;
    }
    return tIntArea;
}

// cutFace::timeIntegratedFaceFlux (cutFace.C:262-389)
template <class CP>
__device__ double faceFlux(const MeshDev& m, int f, const d3& n, double D, double Un0, double dt, double phi,
                           double magSf, int& err)
{
    if (fabs(phi) <= SV_TSMALL) return 0.0;
    d3 fp[CP::MAXFV];
    const int nv = loadFace<CP>(m, f, fp, err);
    const bool flat = __ldg(m.flat + f) > (1.0 - SV_TSMALL);
    d3 c, a, ipd[CP::MAXIP];
    int nip;
    if (fabs(Un0 * dt) > SV_TSMALL) {
        double pTimes[CP::MAXFV];
        for (int i = 0; i < nv; ++i) {
            const double t = (dot(fp[i], n) + D) / Un0;
            pTimes[i] = fabs(t) < SV_TSMALL ? 0.0 : t;
        }
        if (flat) return phi / magSf * timeIntegratedArea<CP>(fp, nv, n, D, pTimes, Un0, dt, magSf, err);
        double dVf = 0.0;
        d3 tri[3];
        double tt[3];
        tri[0] = ld3(m.Cf, f);
        const double t0 = (dot(tri[0], n) + D) / Un0;
        tt[0] = fabs(t0) < SV_TSMALL ? 0.0 : t0;
        for (int pi = 0; pi < nv; ++pi) {
            const int pn = (pi + 1 == nv) ? 0 : pi + 1;
            tri[1] = fp[pi];
            tt[1] = pTimes[pi];
            tri[2] = fp[pn];
            tt[2] = pTimes[pn];
            const double magSfTri = mag(0.5 * cross(tri[1] - tri[0], tri[2] - tri[0]));
            const double phiTri = phi * magSfTri / magSf;
            dVf += phiTri / magSfTri * timeIntegratedArea<CP>(tri, 3, n, D, tt, Un0, dt, magSfTri, err);
        }
        return dVf;
    }
    if (flat) {
        clipFace<CP>(fp, nv, n, D, c, a, ipd, nip, err);
        const double alphaf = mag(a) / magSf;
        return (phi * dt * alphaf);
    }
    d3 tri[3];
    tri[0] = ld3(m.Cf, f);
    double dVf = 0.0;
    for (int pi = 0; pi < nv; ++pi) {
        const int pn = (pi + 1 == nv) ? 0 : pi + 1;
        tri[1] = fp[pi];
        tri[2] = fp[pn];
        const double magSfTri = mag(0.5 * cross(tri[1] - tri[0], tri[2] - tri[0]));
        const double phiTri = phi * magSfTri / magSf;
        clipFace<CP>(tri, 3, n, D, c, a, ipd, nip, err);
        const double alphafTri = mag(a) / magSfTri;
        dVf += (phiTri * dt * alphafTri);
    }
    return dVf;
}

struct SubCellOut {
    int status;
    double VOF, subVol;
    d3 iC, iS;  // interface centre / area vector
    // optional: the interface edge points of a cut cell (cutCell::interfaceEdges_, flattened), for surface extraction
    d3* epOut = nullptr;
    int nEp = 0;
    // optional: cutCell::subCellCentre_ (calcSubCellCentreAndVolume, cutCell.C:103-137), only consumed by subCellFaces()
    bool wantCentre = false;
    d3 subCentre;
};

// face::reverseFace keeps vertex 0 and reverses the rest (cutCell.C:189,211,232)
__device__ __forceinline__ void reversePoly(d3* fp, int nv)
{
    for (int i = 1, j = nv - 1; i < j; ++i, --j) {
        const d3 t = fp[i];
        fp[i] = fp[j];
        fp[j] = t;
    }
}

// cutCell::calcSubCell (cutCell.C:343-542) + calcInterfaceCentreAndArea (:37-100)
// + calcSubCellCentreAndVolume (:103-137).  With split==true the local polyhedron of
// getLocalPointFieldAndFaceList (:140-236) is enumerated on the fly: flat faces as they are
// (reversed when the cell is the neighbour), warped faces as triangle fans about Cf.
template <class CP>
__device__ __noinline__ void subCell(const MeshDev& m, int cell, const d3& n, double D, bool split, SubCellOut& out, int& err)
{
    d3 cfc[CP::MAXCF + 1], cfa[CP::MAXCF + 1];
    d3 ep[CP::MAXEP];
    unsigned char epn[CP::MAXCF];
    int nCut = 0, nEdges = 0, nEp = 0;
    bool fullySubmerged = true, fullyEmpty = true;
    int nSubmergedFaces = 0;

    auto account = [&](int st, const d3& c, const d3& a, const d3* ip, int nip) {
        if (st == 0) {
            if (nCut < CP::MAXCF) {
                cfc[nCut] = c;
                cfa[nCut] = a;
                nCut++;
            } else err |= SVERR_CELL_FACES;
            if (nEdges < CP::MAXCF && nEp + nip <= CP::MAXEP) {
                for (int q = 0; q < nip; ++q) ep[nEp++] = ip[q];
                epn[nEdges++] = (unsigned char)nip;
            } else err |= SVERR_CELL_FACES;
            fullySubmerged = false;
            fullyEmpty = false;
        } else if (st == -1) {
            if (nCut < CP::MAXCF) {
                cfc[nCut] = c;
                cfa[nCut] = a;
                nCut++;
            } else err |= SVERR_CELL_FACES;
            fullyEmpty = false;
            nSubmergedFaces++;
        } else {
            fullySubmerged = false;
        }
    };

    const int c0 = __ldg(m.cellOff + cell), c1 = __ldg(m.cellOff + cell + 1);
    d3 fp[CP::MAXFV], fc, fa, ip[CP::MAXIP];
    int nip;
    for (int k = c0; k < c1; ++k) {
        const int f = __ldg(m.cellFaces + k);
        const int nv = loadFace<CP>(m, f, fp, err);
        if (!split) {
            const int st = clipFace<CP>(fp, nv, n, D, fc, fa, ip, nip, err);
            account(st, fc, fa, ip, nip);
        } else {
            const bool own = (__ldg(m.owner + f) == cell);
            if (__ldg(m.flat + f) > (1.0 - SV_TSMALL)) {
                if (!own) reversePoly(fp, nv);
                const int st = clipFace<CP>(fp, nv, n, D, fc, fa, ip, nip, err);
                account(st, fc, fa, ip, nip);
            } else {
                d3 tri[3];
                const d3 ctr = ld3(m.Cf, f);
                for (int pi = 0; pi < nv; ++pi) {
                    const int pn = (pi + 1 == nv) ? 0 : pi + 1;
                    tri[0] = ctr;
                    tri[1] = own ? fp[pi] : fp[pn];  // reverseFace of (c, p_i, p_{i+1}) = (c, p_{i+1}, p_i)
                    tri[2] = own ? fp[pn] : fp[pi];
                    const int st = clipFace<CP>(tri, 3, n, D, fc, fa, ip, nip, err);
                    account(st, fc, fa, ip, nip);
                }
            }
        }
    }

    out.iC = zero3();
    out.iS = zero3();
    if (!fullySubmerged && !fullyEmpty) {
        out.status = 0;
        // calcInterfaceCentreAndArea
        d3 fC = zero3();
        for (int q = 0; q < nEp; ++q) fC += ep[q];
        if (nEp > 0) fC /= double(nEp);
        d3 sumN = zero3(), sumAc = zero3();
        double sumA = 0.0;
        int base = 0;
        for (int e = 0; e < nEdges; ++e) {
            const int np = epn[e];
            for (int pi = 0; pi < np - 1; ++pi) {
                const d3 p0 = ep[base + pi], nx = ep[base + pi + 1];
                const d3 c = p0 + nx + fC;
                const d3 nn = cross(nx - p0, fC - p0);
                const double a = mag(nn);
                sumN += sgn(dot(nn, sumN)) * nn;
                sumA += a;
                sumAc += a * c;
            }
            base += np;
        }
        d3 iC, iS;
        if (sumA < SV_ROOTVSMALL) {
            iC = fC;
            iS = zero3();
        } else {
            iC = (1.0 / 3.0) * sumAc / sumA;
            iS = 0.5 * sumN;
        }
        // the reference tests against subCellCentre_, which is still (0,0,0) here (SURVEY 8a' item 24)
        if (dot(iS, iC - zero3()) < 0.0) iS = iS * (-1.0);
        out.iC = iC;
        out.iS = iS;
        if (mag(iS) < SV_TSMALL) {
            if (nSubmergedFaces == 0) {
                out.status = 1;
                out.subVol = 0.0;
                out.VOF = 0.0;
            } else {
                out.status = -1;
                out.subVol = __ldg(m.V + cell);
                out.VOF = 1.0;
            }
            return;
        }
        if (out.epOut) {
            for (int q = 0; q < nEp; ++q) out.epOut[q] = ep[q];
            out.nEp = nEp;
        }
        cfc[nCut] = iC;
        cfa[nCut] = iS;
        nCut++;
        // calcSubCellCentreAndVolume (centre itself is not consumed on this path)
        d3 cEst = zero3();
        for (int q = 0; q < nCut; ++q) cEst += cfc[q];
        cEst /= double(nCut);
        double vol = 0.0;
        for (int q = 0; q < nCut; ++q) vol += dmax(fabs(dot(cfa[q], cfc[q] - cEst)), SV_VSMALL);
        if (out.wantCentre) {
            d3 sc = zero3();
            for (int q = 0; q < nCut; ++q) {
                const double pyr3Vol = dmax(fabs(dot(cfa[q], cfc[q] - cEst)), SV_VSMALL);
                const d3 pc = 0.75 * cfc[q] + 0.25 * cEst;
                sc += pyr3Vol * pc;
            }
            sc /= vol;
            out.subCentre = sc;
        }
        vol /= 3.0;
        out.subVol = vol;
        out.VOF = vol / __ldg(m.V + cell);
    } else if (fullyEmpty) {
        out.status = 1;
        out.subVol = 0.0;
        out.VOF = 0.0;
    } else {
        out.status = -1;
        out.subVol = __ldg(m.V + cell);
        out.VOF = 1.0;
    }
}

struct PlicOut {
    int status;
    bool wrote;  // the reference leaves D/C/S untouched on the |n| < TSMALL early return
    double D;
    d3 C, S;
};

// cutCell::findSignedDistance (cutCell.C:611-799)
template <class CP>
__device__ void signedDistance(const MeshDev& m, int cell, double alphaI, const d3& n, bool split, PlicOut& po, int& err)
{
    po.wrote = false;
    po.D = 0;
    po.C = zero3();
    po.S = zero3();
    if (mag(n) < SV_TSMALL) {
        po.status = int(sgn(0.5 - alphaI));
        return;
    }
    double vd[CP::MAXCP];
    int nP = 0;
    {
        const int p0 = __ldg(m.cellPtOff + cell), p1 = __ldg(m.cellPtOff + cell + 1);
        for (int k = p0; k < p1; ++k) {
            if (nP < CP::MAXCP) vd[nP++] = -dot(n, ld3(m.points, __ldg(m.cellPts + k)));
            else err |= SVERR_CELL_POINTS;
        }
        if (split) {  // one appended point (the face centre) per warped face
            const int c0 = __ldg(m.cellOff + cell), c1 = __ldg(m.cellOff + cell + 1);
            for (int k = c0; k < c1; ++k) {
                const int f = __ldg(m.cellFaces + k);
                if (!(__ldg(m.flat + f) > (1.0 - SV_TSMALL))) {
                    if (nP < CP::MAXCP) vd[nP++] = -dot(n, ld3(m.Cf, f));
                    else err |= SVERR_CELL_POINTS;
                }
            }
        }
    }
    sortDescending(vd, nP);

    double lowDistance = vd[0], upDistance = vd[nP - 1];
    int lowLabel = 0, upLabel = nP - 1;
    double lowAlpha = 0.0, upAlpha = 1.0;
    SubCellOut sc;
    while ((upLabel - lowLabel) > 1) {
        const double midLabel = round(0.5 * (upLabel + lowLabel));  // a scalar in the reference (:689-693)
        const double midDistance = vd[int(midLabel)];
        subCell<CP>(m, cell, n, midDistance, split, sc, err);
        const double midAlpha = sc.VOF;
        if (fabs(midAlpha - alphaI) < SV_TSMALL) {
            po.status = sc.status;
            po.wrote = true;
            po.D = midDistance;
            po.C = sc.iC;
            po.S = sc.iS;
            return;
        }
        if (midAlpha > alphaI) {
            upLabel = int(midLabel);
            upDistance = midDistance;
            upAlpha = midAlpha;
        } else {
            lowLabel = int(midLabel);
            lowDistance = midDistance;
            lowAlpha = midAlpha;
        }
    }
    if (fabs(lowDistance - upDistance) < SV_TSMALL) {
        const double midD = 0.5 * (lowDistance + upDistance);
        subCell<CP>(m, cell, n, midD, split, sc, err);
        po.status = sc.status;
        po.wrote = true;
        po.D = midD;
        po.C = sc.iC;
        po.S = sc.iS;
        return;
    }
    const double alphaPrismatoid = upAlpha - lowAlpha;
    const double deltaDistance = (upDistance - lowDistance) / 3.0;
    const double distanceOneThird = lowDistance + deltaDistance;
    subCell<CP>(m, cell, n, distanceOneThird, split, sc, err);
    const double alphaOneThird = sc.VOF - lowAlpha;
    const double distanceTwoThirds = lowDistance + 2.0 * deltaDistance;
    subCell<CP>(m, cell, n, distanceTwoThirds, split, sc, err);
    const double alphaTwoThirds = sc.VOF - lowAlpha;

    const double a = 13.5 * alphaOneThird - 13.5 * alphaTwoThirds + 4.5 * alphaPrismatoid;
    const double b = -22.5 * alphaOneThird + 18.0 * alphaTwoThirds - 4.5 * alphaPrismatoid;
    const double c = 9.0 * alphaOneThird - 4.5 * alphaTwoThirds + 1.0 * alphaPrismatoid;
    const double d = lowAlpha - alphaI;
    double lambda = 0.5;
    for (int iter = 0; iter < 100; ++iter) {
        const double func = a * (lambda * (lambda * lambda)) + b * (lambda * lambda) + c * lambda + d;
        const double funcPrime = 3.0 * a * (lambda * lambda) + 2.0 * b * lambda + c;
        const double lambdaNew = lambda - (func / funcPrime);
        if (fabs(lambdaNew - lambda) < SV_TSMALL) break;
        lambda = lambdaNew;
    }
    const double distance0 = lowDistance - lambda * (lowDistance - upDistance);
    subCell<CP>(m, cell, n, distance0, split, sc, err);
    po.status = sc.status;
    po.wrote = true;
    po.D = distance0;
    po.C = sc.iC;
    po.S = sc.iS;
}

// ---- Foam::LUDecompose/LUBacksubstitute (OF, recalled): Crout LU, implicit-scaling pivoting ----
__device__ __forceinline__ void luSolve4(double A[4][4], double b[4], int mdim)
{
    int pivot[4];
    double vv[4];
    for (int i = 0; i < mdim; ++i) {
        double largest = 0.0, t;
        for (int j = 0; j < mdim; ++j)
            if ((t = fabs(A[i][j])) > largest) largest = t;
        if (largest == 0.0) largest = SV_SMALL;
        vv[i] = 1.0 / largest;
    }
    for (int j = 0; j < mdim; ++j) {
        for (int i = 0; i < j; ++i) {
            double sum = A[i][j];
            for (int k = 0; k < i; ++k) sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
        }
        int iMax = 0;
        double largest = 0.0;
        for (int i = j; i < mdim; ++i) {
            double sum = A[i][j];
            for (int k = 0; k < j; ++k) sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
            double t;
            if ((t = vv[i] * fabs(sum)) >= largest) {
                largest = t;
                iMax = i;
            }
        }
        pivot[j] = iMax;
        if (j != iMax) {
            for (int k = 0; k < mdim; ++k) {
                const double t = A[j][k];
                A[j][k] = A[iMax][k];
                A[iMax][k] = t;
            }
            vv[iMax] = vv[j];
        }
        if (A[j][j] == 0.0) A[j][j] = SV_SMALL;
        if (j != mdim - 1) {
            const double rDiag = 1.0 / A[j][j];
            for (int i = j + 1; i < mdim; ++i) A[i][j] *= rDiag;
        }
    }
    int ii = 0;
    for (int i = 0; i < mdim; ++i) {
        const int ip = pivot[i];
        double sum = b[ip];
        b[ip] = b[i];
        if (ii != 0) {
            for (int j = ii - 1; j < i; ++j) sum -= A[i][j] * b[j];
        } else if (sum != 0.0) {
            ii = i + 1;
        }
        b[i] = sum;
    }
    for (int i = mdim - 1; i >= 0; --i) {
        double sum = b[i];
        for (int j = i + 1; j < mdim; ++j) sum -= A[i][j] * b[j];
        b[i] = sum / A[i][i];
    }
}

}  // namespace svof
