// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Plain-array polyhedral mesh + the OpenFOAM mesh services the SimPLIC path
// calls.  OpenFOAM v2312 is not vendored in the reference tree, so the
// formulas below are restated from its published source (marked "OF,
// recalled") and anchored on the reference's call sites:
//   mesh_.faceCentres()/faceAreas()   cutFace.C:318, reconstruction.C:412-413
//   mesh_.cellCentres()/cellVolumes() cutCell.C:357-358
//   mesh_.cells()[c]                  cutCell.C:427, advection.C:103,131
//   mesh_.cellPoints(c)               cutCell.C:637,656
//   face flatness                     reconstruction.C:408-473
//   pointCells / boundary pointFaces  volPointInterpolation (advection.C:91)
//   tetBasePtIs                       interpolationCellPoint (advection.C:126)
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/svof.h"
#include "ora_vec.hpp"

namespace ora {

struct Csr {
    std::vector<label> off, idx;
    label size(label i) const { return off[i + 1] - off[i]; }
    const label* row(label i) const { return idx.data() + off[i]; }
};

struct Mesh {
    label nPoints = 0, nFaces = 0, nInternalFaces = 0, nCells = 0;
    std::vector<point> points;
    Csr faces;  // face -> point labels
    std::vector<label> owner, neighbour;
    std::vector<svof_patch> patches;

    // derived geometry
    std::vector<vec> Cf, Sf, C;
    std::vector<scalar> magSf, V, faceFlatness;
    scalar flatMin = 1, flatMax = 1, flatAvg = 1;

    // derived addressing
    Csr cells;       // primitiveMesh::cells(): owned faces ascending, then neighbour-side faces ascending
    Csr cellPoints;  // ascending point label (order is irrelevant to results, cutCell.C:664-670)
    Csr pointCells;  // ascending cell label (primitiveMesh::calcPointCells)
    Csr pointBFaces; // point -> boundary faces (index f - nInternalFaces), ascending
    std::vector<label> patchID;       // per boundary face
    std::vector<char> isPatchFace;    // per boundary face: patch neither empty nor coupled
    std::vector<char> isPatchPoint;   // per point: on such a face
    std::vector<label> tetBasePt;     // per face (polyMesh::tetBasePtIs)

    label nBoundaryFaces() const { return nFaces - nInternalFaces; }
    bool isInternalFace(label f) const { return f < nInternalFaces; }

    void build(const svof_mesh& m);
    // mesh.moving(): new point positions, same topology -- geometry, flatness (reconstruction.C:643-647) and tet base points follow
    void movePoints(const double* pts, const double* hCf, const double* hSf, const double* hC, const double* hV);

   private:
    void calcFaceCentresAndAreas();
    void calcCellCentresAndVols();
    void calcFaceFlatness();
    void calcTetBasePts();
};

// ---- primitiveMeshTools::faceCentresAndAreas (OF, recalled) ------------------
inline void faceCentreAndArea(const point* p, const label* f, label n, vec& fc, vec& fa)
{
    if (n == 3) {
        fc = (1.0 / 3.0) * (p[f[0]] + p[f[1]] + p[f[2]]);
        fa = 0.5 * ((p[f[1]] - p[f[0]]) ^ (p[f[2]] - p[f[0]]));
        return;
    }
    vec sumN, sumAc;
    scalar sumA = 0.0;
    vec fCentre = p[f[0]];
    for (label pi = 1; pi < n; ++pi) fCentre += p[f[pi]];
    fCentre /= scalar(n);
    for (label pi = 0; pi < n; ++pi) {
        const point& nextPoint = p[f[(pi == n - 1) ? 0 : pi + 1]];
        const point& thisPoint = p[f[pi]];
        vec c = thisPoint + nextPoint + fCentre;
        vec nrm = (nextPoint - thisPoint) ^ (fCentre - thisPoint);
        scalar a = mag(nrm);
        sumN += nrm;
        sumA += a;
        sumAc += a * c;
    }
    if (sumA < ROOTVSMALL) {
        fc = fCentre;
        fa = vec();
    } else {
        fc = (1.0 / 3.0) * sumAc / sumA;
        fa = 0.5 * sumN;
    }
}

inline void Mesh::calcFaceCentresAndAreas()
{
    Cf.resize(nFaces);
    Sf.resize(nFaces);
    magSf.resize(nFaces);
    for (label f = 0; f < nFaces; ++f) {
        faceCentreAndArea(points.data(), faces.row(f), faces.size(f), Cf[f], Sf[f]);
    }
}

// ---- primitiveMeshTools::cellCentresAndVols (OF, recalled) -------------------
inline void Mesh::calcCellCentresAndVols()
{
    C.assign(nCells, vec());
    V.assign(nCells, 0.0);
    std::vector<vec> cEst(nCells);
    std::vector<label> nCellFaces(nCells, 0);
    for (label f = 0; f < nFaces; ++f) {
        cEst[owner[f]] += Cf[f];
        ++nCellFaces[owner[f]];
    }
    for (label f = 0; f < nInternalFaces; ++f) {
        cEst[neighbour[f]] += Cf[f];
        ++nCellFaces[neighbour[f]];
    }
    for (label c = 0; c < nCells; ++c) cEst[c] /= scalar(nCellFaces[c]);
    for (label f = 0; f < nFaces; ++f) {
        const label c = owner[f];
        scalar pyr3Vol = Sf[f] & (Cf[f] - cEst[c]);
        vec pc = (3.0 / 4.0) * Cf[f] + (1.0 / 4.0) * cEst[c];
        C[c] += pyr3Vol * pc;
        V[c] += pyr3Vol;
    }
    for (label f = 0; f < nInternalFaces; ++f) {
        const label c = neighbour[f];
        scalar pyr3Vol = Sf[f] & (cEst[c] - Cf[f]);
        vec pc = (3.0 / 4.0) * Cf[f] + (1.0 / 4.0) * cEst[c];
        C[c] += pyr3Vol * pc;
        V[c] += pyr3Vol;
    }
    for (label c = 0; c < nCells; ++c) {
        if (mag(V[c]) > VSMALL) {
            C[c] /= V[c];
        } else {
            C[c] = cEst[c];
        }
    }
    for (label c = 0; c < nCells; ++c) V[c] *= (1.0 / 3.0);
}

// ---- reconstruction::updateFaceFlatness (reconstruction.C:408-447) ----------
inline void Mesh::calcFaceFlatness()
{
    faceFlatness.assign(nFaces, 1.0);
    scalar sumFA = 0, sumA = 0;
    flatMin = VGREAT;
    flatMax = -VGREAT;
    for (label f = 0; f < nFaces; ++f) {
        const label n = faces.size(f);
        const label* fa = faces.row(f);
        if (n > 3 && magSf[f] > ROOTVSMALL) {
            const point& fc = Cf[f];
            scalar sA = 0.0;
            for (label pi = 0; pi < n; ++pi) {
                const point& thisPoint = points[fa[pi]];
                const point& nextPoint = points[fa[(pi + 1) % n]];
                vec nrm = 0.5 * ((nextPoint - thisPoint) ^ (fc - thisPoint));
                sA += mag(nrm);
            }
            faceFlatness[f] = magSf[f] / (sA + ROOTVSMALL);
        } else {
            faceFlatness[f] = 1.0;
        }
        flatMin = smin(flatMin, faceFlatness[f]);
        flatMax = smax(flatMax, faceFlatness[f]);
        sumFA += faceFlatness[f] * magSf[f];
        sumA += magSf[f];
    }
    flatAvg = sumFA / sumA;
}

// ---- polyMeshTetDecomposition::findFaceBasePts (OF, recalled) -----------------
// First face vertex whose fan tets (owner and neighbour side) all have
// tetrahedron::quality() > 1e-9; vertex 0 on any sane mesh.
inline scalar tetQuality(const point& a, const point& b, const point& c, const point& d)
{
    // tetrahedron::mag(): (1/6) ((b-a)^(c-a)) & (d-a)
    const scalar vol = (1.0 / 6.0) * (((b - a) ^ (c - a)) & (d - a));
    // tetrahedron::circumRadius()
    const vec ea = b - a, eb = c - a, ec = d - a;
    const scalar lambda = magSqr(ec) - (ea & ec);
    const scalar mu = magSqr(eb) - (ea & eb);
    const vec ba = eb ^ ea, ca = ec ^ ea;
    const vec num = lambda * ba - mu * ca;
    const scalar denom = (ec & ba);
    scalar R = GREAT;
    if (mag(denom) >= ROOTVSMALL) R = mag(0.5 * (ea + num / denom));
    return vol / ((8.0 / (9.0 * std::sqrt(3.0))) * pow3(smin(R, GREAT)) + ROOTVSMALL);
}

inline void Mesh::calcTetBasePts()
{
    tetBasePt.assign(nFaces, 0);
    const scalar tol = 1e-9;  // polyMeshTetDecomposition::minTetQuality
    for (label f = 0; f < nFaces; ++f) {
        const label n = faces.size(f);
        const label* fp = faces.row(f);
        const bool internal = f < nInternalFaces;
        const point& oCc = C[owner[f]];
        label found = -1;
        for (label base = 0; base < n && found < 0; ++base) {
            scalar minQ = VGREAT;
            const point& pb = points[fp[base]];
            for (label t = 1; t < n - 1; ++t) {
                const label ia = (t + base) % n;
                const label ib = (ia + 1) % n;
                scalar q = tetQuality(oCc, pb, points[fp[ia]], points[fp[ib]]);
                if (internal) {
                    const point& nCc = C[neighbour[f]];
                    q = smin(q, tetQuality(nCc, pb, points[fp[ib]], points[fp[ia]]));
                }
                if (q < minQ) minQ = q;
            }
            if (minQ > tol) found = base;
        }
        // tetIndices::faceTriIs falls back to 0 when no base point qualifies
        tetBasePt[f] = (found < 0) ? 0 : found;
    }
}

inline void Mesh::movePoints(const double* pts, const double* hCf, const double* hSf, const double* hC, const double* hV)
{
    for (label i = 0; i < nPoints; ++i) points[i] = point(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    if (hCf && hSf) {
        for (label f = 0; f < nFaces; ++f) {
            Cf[f] = vec(hCf[3 * f], hCf[3 * f + 1], hCf[3 * f + 2]);
            Sf[f] = vec(hSf[3 * f], hSf[3 * f + 1], hSf[3 * f + 2]);
        }
    } else {
        calcFaceCentresAndAreas();
    }
    for (label f = 0; f < nFaces; ++f) magSf[f] = mag(Sf[f]);
    if (hC && hV) {
        V.assign(hV, hV + nCells);
        for (label c = 0; c < nCells; ++c) C[c] = vec(hC[3 * c], hC[3 * c + 1], hC[3 * c + 2]);
    } else {
        calcCellCentresAndVols();
    }
    calcFaceFlatness();
    calcTetBasePts();
}

inline void Mesh::build(const svof_mesh& m)
{
    if (!m.points || !m.face_offsets || !m.face_points || !m.owner || (m.n_internal_faces > 0 && !m.neighbour))
        throw std::invalid_argument("svof_mesh: null connectivity pointer");
    nPoints = m.n_points;
    nFaces = m.n_faces;
    nInternalFaces = m.n_internal_faces;
    nCells = m.n_cells;
    if (nPoints <= 0 || nFaces <= 0 || nCells <= 0 || nInternalFaces < 0 || nInternalFaces > nFaces)
        throw std::invalid_argument("svof_mesh: bad sizes");
    points.resize(nPoints);
    for (label i = 0; i < nPoints; ++i) points[i] = point(m.points[3 * i], m.points[3 * i + 1], m.points[3 * i + 2]);
    faces.off.assign(m.face_offsets, m.face_offsets + nFaces + 1);
    faces.idx.assign(m.face_points, m.face_points + faces.off[nFaces]);
    owner.assign(m.owner, m.owner + nFaces);
    neighbour.assign(m.neighbour, m.neighbour + nInternalFaces);
    patches.assign(m.patches, m.patches + m.n_patches);
    for (label f = 0; f < nFaces; ++f) {
        if (owner[f] < 0 || owner[f] >= nCells) throw std::invalid_argument("svof_mesh: owner out of range");
        if (faces.size(f) < 3) throw std::invalid_argument("svof_mesh: face with < 3 points");
        if (f < nInternalFaces && (neighbour[f] < 0 || neighbour[f] >= nCells))
            throw std::invalid_argument("svof_mesh: neighbour out of range");
    }
    for (label i : faces.idx)
        if (i < 0 || i >= nPoints) throw std::invalid_argument("svof_mesh: point label out of range");

    // patch table: contiguous cover of the boundary faces
    const label nBF = nBoundaryFaces();
    patchID.assign(nBF, -1);
    isPatchFace.assign(nBF, 0);
    label expect = nInternalFaces;
    for (size_t pi = 0; pi < patches.size(); ++pi) {
        const svof_patch& p = patches[pi];
        if (p.start != expect || p.size < 0 || p.start + p.size > nFaces)
            throw std::invalid_argument("svof_mesh: patches must tile the boundary faces in order");
        for (label k = 0; k < p.size; ++k) {
            patchID[p.start - nInternalFaces + k] = label(pi);
            isPatchFace[p.start - nInternalFaces + k] = (p.kind == SVOF_PATCH_GENERIC);
        }
        expect += p.size;
    }
    if (expect != nFaces) throw std::invalid_argument("svof_mesh: patches do not cover all boundary faces");

    // geometry
    if (m.Cf && m.Sf) {
        Cf.resize(nFaces);
        Sf.resize(nFaces);
        for (label f = 0; f < nFaces; ++f) {
            Cf[f] = vec(m.Cf[3 * f], m.Cf[3 * f + 1], m.Cf[3 * f + 2]);
            Sf[f] = vec(m.Sf[3 * f], m.Sf[3 * f + 1], m.Sf[3 * f + 2]);
        }
        magSf.resize(nFaces);
    } else {
        calcFaceCentresAndAreas();
    }
    for (label f = 0; f < nFaces; ++f) magSf[f] = mag(Sf[f]);
    if (m.C && m.V) {
        C.resize(nCells);
        V.assign(m.V, m.V + nCells);
        for (label c = 0; c < nCells; ++c) C[c] = vec(m.C[3 * c], m.C[3 * c + 1], m.C[3 * c + 2]);
    } else {
        calcCellCentresAndVols();
    }

    // cells(): primitiveMesh::calcCells -- owner loop then neighbour loop
    cells.off.assign(nCells + 1, 0);
    for (label f = 0; f < nFaces; ++f) cells.off[owner[f] + 1]++;
    for (label f = 0; f < nInternalFaces; ++f) cells.off[neighbour[f] + 1]++;
    for (label c = 0; c < nCells; ++c) cells.off[c + 1] += cells.off[c];
    cells.idx.resize(cells.off[nCells]);
    {
        std::vector<label> fill(cells.off.begin(), cells.off.end() - 1);
        for (label f = 0; f < nFaces; ++f) cells.idx[fill[owner[f]]++] = f;
        for (label f = 0; f < nInternalFaces; ++f) cells.idx[fill[neighbour[f]]++] = f;
    }

    // cellPoints (ascending) and pointCells (ascending)
    cellPoints.off.assign(nCells + 1, 0);
    {
        std::vector<label> tmp;
        auto gather = [&](label c) {
            tmp.clear();
            for (label k = 0; k < cells.size(c); ++k) {
                const label f = cells.row(c)[k];
                tmp.insert(tmp.end(), faces.row(f), faces.row(f) + faces.size(f));
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        };
        for (label c = 0; c < nCells; ++c) {
            gather(c);
            cellPoints.off[c + 1] = cellPoints.off[c] + label(tmp.size());
        }
        cellPoints.idx.resize(cellPoints.off[nCells]);
        for (label c = 0; c < nCells; ++c) {
            gather(c);
            std::copy(tmp.begin(), tmp.end(), cellPoints.idx.begin() + cellPoints.off[c]);
        }
    }
    pointCells.off.assign(nPoints + 1, 0);
    for (label i : cellPoints.idx) pointCells.off[i + 1]++;
    for (label p = 0; p < nPoints; ++p) pointCells.off[p + 1] += pointCells.off[p];
    pointCells.idx.resize(pointCells.off[nPoints]);
    {
        std::vector<label> fill(pointCells.off.begin(), pointCells.off.end() - 1);
        for (label c = 0; c < nCells; ++c)
            for (label k = 0; k < cellPoints.size(c); ++k) pointCells.idx[fill[cellPoints.row(c)[k]]++] = c;
    }

    // boundary point -> boundary faces (PrimitivePatch::pointFaces order: ascending face)
    pointBFaces.off.assign(nPoints + 1, 0);
    isPatchPoint.assign(nPoints, 0);
    for (label bf = 0; bf < nBF; ++bf) {
        const label f = nInternalFaces + bf;
        for (label k = 0; k < faces.size(f); ++k) {
            pointBFaces.off[faces.row(f)[k] + 1]++;
            if (isPatchFace[bf]) isPatchPoint[faces.row(f)[k]] = 1;
        }
    }
    for (label p = 0; p < nPoints; ++p) pointBFaces.off[p + 1] += pointBFaces.off[p];
    pointBFaces.idx.resize(pointBFaces.off[nPoints]);
    {
        std::vector<label> fill(pointBFaces.off.begin(), pointBFaces.off.end() - 1);
        for (label bf = 0; bf < nBF; ++bf) {
            const label f = nInternalFaces + bf;
            for (label k = 0; k < faces.size(f); ++k) pointBFaces.idx[fill[faces.row(f)[k]]++] = bf;
        }
    }

    calcFaceFlatness();
    calcTetBasePts();
}

}  // namespace ora
