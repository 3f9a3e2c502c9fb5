// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Restatement of the reference's L1 geometry classes, statement by statement:
//   cutFace  <- src/SimPLIC/cut/cutFace/cutFace.C
//   cutCell  <- src/SimPLIC/cut/cutCell/cutCell.C
// Member names follow the reference so the two can be read side by side; the
// containers are std::vector instead of DynamicList, the arithmetic and its
// order are the reference's.  OpenFOAM services (face::centre,
// face::areaNormal, triFace::mag, sortedOrder) are restated from OF v2312
// ("OF, recalled").
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "ora_mesh.hpp"

namespace ora {

static const scalar TSMALL = 10.0 * SMALL;  // cutFace.C:154, cutCell.C:355,618

// face::centre(points) (OF, recalled)
inline point faceCentreOF(const std::vector<point>& p)
{
    const label nPoints = label(p.size());
    if (nPoints == 3) return (1.0 / 3.0) * (p[0] + p[1] + p[2]);
    point centrePoint;
    for (label pI = 0; pI < nPoints; ++pI) centrePoint += p[pI];
    centrePoint /= scalar(nPoints);
    scalar sumA = 0;
    vec sumAc;
    for (label pI = 0; pI < nPoints; ++pI) {
        const point& nextPoint = p[(pI + 1) % nPoints];
        const vec ttc(p[pI] + nextPoint + centrePoint);                       // 3*triangle centre
        const scalar ta = mag((p[pI] - centrePoint) ^ (nextPoint - centrePoint));  // 2*triangle area
        sumA += ta;
        sumAc += ta * ttc;
    }
    if (sumA > VSMALL) return sumAc / (3.0 * sumA);
    return centrePoint;
}

// face::areaNormal(points) (OF, recalled)
inline vec faceAreaNormalOF(const std::vector<point>& p)
{
    const label nPoints = label(p.size());
    if (nPoints == 3) return 0.5 * ((p[1] - p[0]) ^ (p[2] - p[0]));
    point centrePoint;
    for (label pI = 0; pI < nPoints; ++pI) centrePoint += p[pI];
    centrePoint /= scalar(nPoints);
    vec n;
    for (label pI = 0; pI < nPoints; ++pI) {
        const point& nextPoint = (pI < nPoints - 1) ? p[pI + 1] : p[0];
        n += 0.5 * ((nextPoint - p[pI]) ^ (centrePoint - p[pI]));  // triPointRef::areaNormal(a,b,c)
    }
    return n;
}

// Foam::sortedOrder: identity permutation + std::stable_sort (ascending)
inline std::vector<label> sortedOrderLess(const std::vector<scalar>& v)
{
    std::vector<label> order(v.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](label a, label b) { return v[a] < v[b]; });
    return order;
}
inline std::vector<label> sortedOrderGreater(const std::vector<scalar>& v)
{
    std::vector<label> order(v.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](label a, label b) { return v[a] > v[b]; });
    return order;
}

// ---------------------------------------------------------------- cutFace ----
class cutFace {
    const Mesh& mesh_;
    point subFaceCentre_;
    vec subFaceArea_;
    std::vector<point> subFacePoints_;
    std::vector<point> interfacePoints_;
    std::vector<scalar> pointDistances_;
    label faceStatus_;

    // cutFace.C:37-96
    void calcSubFaceCentreAndArea()
    {
        const label nPoints = label(subFacePoints_.size());
        if (nPoints == 3) {
            subFaceCentre_ = (1.0 / 3.0) * (subFacePoints_[0] + subFacePoints_[1] + subFacePoints_[2]);
            subFaceArea_ = 0.5 * ((subFacePoints_[1] - subFacePoints_[0]) ^ (subFacePoints_[2] - subFacePoints_[0]));
        } else {
            vec sumN;
            scalar sumA = 0.0;
            vec sumAc;
            point fCentre(subFacePoints_[0]);
            for (label pi = 1; pi < nPoints; pi++) fCentre += subFacePoints_[pi];
            fCentre /= scalar(nPoints);
            for (label pi = 0; pi < nPoints; pi++) {
                const point& nextPoint = subFacePoints_[(pi + 1) % nPoints];
                vec c(subFacePoints_[pi] + nextPoint + fCentre);
                vec n((nextPoint - subFacePoints_[pi]) ^ (fCentre - subFacePoints_[pi]));
                scalar a = mag(n);
                sumN += n;
                sumA += a;
                sumAc += a * c;
            }
            if (sumA < ROOTVSMALL) {
                subFaceCentre_ = fCentre;
                subFaceArea_ = vec();
            } else {
                subFaceCentre_ = (1.0 / 3.0) * sumAc / sumA;
                subFaceArea_ = 0.5 * sumN;
            }
        }
    }

   public:
    explicit cutFace(const Mesh& mesh) : mesh_(mesh), faceStatus_(-1) { clearStorage(); }

    void clearStorage()  // cutFace.H:187-195
    {
        subFaceCentre_ = point();
        subFaceArea_ = vec();
        subFacePoints_.clear();
        interfacePoints_.clear();
        pointDistances_.clear();
        faceStatus_ = -1;
    }

    std::vector<point> facePoints(label faceI) const
    {
        std::vector<point> fPts(mesh_.faces.size(faceI));
        for (size_t i = 0; i < fPts.size(); ++i) fPts[i] = mesh_.points[mesh_.faces.row(faceI)[i]];
        return fPts;
    }

    // cutFace.C:122-133
    label calcSubFace(label faceI, const vec& normal, scalar distance)
    {
        return calcSubFace(facePoints(faceI), normal, distance);
    }

    // cutFace.C:136-259
    label calcSubFace(const std::vector<point>& fPts, const vec& normal, scalar distance)
    {
        clearStorage();
        const label fsize = label(fPts.size());
        label nSubmergedPoints = 0;
        label firstSubmergedPoint = -1;

        for (label i = 0; i < fsize; ++i) {
            scalar distanceI = (fPts[i] & normal) + distance;
            if (mag(distanceI) < TSMALL) distanceI += sign(distanceI) * TSMALL;
            pointDistances_.push_back(distanceI);
            if (distanceI < 0.0) {
                nSubmergedPoints++;
                if (firstSubmergedPoint == -1) firstSubmergedPoint = i;
            }
        }

        if (nSubmergedPoints == fsize) {
            faceStatus_ = -1;
            subFaceCentre_ = faceCentreOF(fPts);
            subFaceArea_ = faceAreaNormalOF(fPts);
            return faceStatus_;
        } else if (nSubmergedPoints == 0) {
            faceStatus_ = 1;
            subFaceCentre_ = point();
            subFaceArea_ = vec();
            return faceStatus_;
        } else {
            faceStatus_ = 0;
            for (label i = firstSubmergedPoint; i < firstSubmergedPoint + fsize; ++i) {
                const label currentId = i % fsize;
                const label nextId = (i + 1) % fsize;
                if (pointDistances_[currentId] < 0) subFacePoints_.push_back(fPts[currentId]);
                if ((pointDistances_[currentId] * pointDistances_[nextId]) < 0) {
                    const scalar weight = pointDistances_[currentId] / (pointDistances_[currentId] - pointDistances_[nextId]);
                    const point cutPoint(fPts[currentId] + weight * (fPts[nextId] - fPts[currentId]));
                    subFacePoints_.push_back(cutPoint);
                    interfacePoints_.push_back(cutPoint);
                }
            }
            if (subFacePoints_.size() >= 3) {
                faceStatus_ = 0;
                calcSubFaceCentreAndArea();
            } else {
                faceStatus_ = -1;
                subFaceCentre_ = faceCentreOF(fPts);
                subFaceArea_ = faceAreaNormalOF(fPts);
            }
            return faceStatus_;
        }
    }

    // cutFace.C:262-389
    scalar timeIntegratedFaceFlux(label faceI, const vec& normal, scalar distance, scalar Un0, scalar dt, scalar phi,
                                  scalar magSf)
    {
        if (mag(phi) <= TSMALL) return 0.0;

        const std::vector<point> fPts(facePoints(faceI));
        const label nPoints = label(fPts.size());

        if (mag(Un0 * dt) > TSMALL) {
            std::vector<scalar> pTimes(nPoints);
            for (label i = 0; i < nPoints; ++i) {
                scalar pTimeI = ((fPts[i] & normal) + distance) / Un0;
                pTimes[i] = mag(pTimeI) < TSMALL ? 0.0 : pTimeI;
            }
            scalar dVf = 0.0;
            if (mesh_.faceFlatness[faceI] > (1.0 - TSMALL)) {
                dVf = phi / magSf * timeIntegratedArea(fPts, normal, distance, pTimes, Un0, dt, magSf);
            } else {
                std::vector<point> fPtsTri(3);
                std::vector<scalar> pTimesTri(3);
                fPtsTri[0] = mesh_.Cf[faceI];
                scalar pTimeTri0 = ((fPtsTri[0] & normal) + distance) / Un0;
                pTimesTri[0] = mag(pTimeTri0) < TSMALL ? 0.0 : pTimeTri0;
                for (label pi = 0; pi < nPoints; ++pi) {
                    fPtsTri[1] = fPts[pi];
                    pTimesTri[1] = pTimes[pi];
                    fPtsTri[2] = fPts[(pi + 1) % nPoints];
                    pTimesTri[2] = pTimes[(pi + 1) % nPoints];
                    const scalar magSfTri = mag(0.5 * ((fPtsTri[1] - fPtsTri[0]) ^ (fPtsTri[2] - fPtsTri[0])));  // triFace::mag
                    const scalar phiTri = phi * magSfTri / magSf;
                    dVf += phiTri / magSfTri * timeIntegratedArea(fPtsTri, normal, distance, pTimesTri, Un0, dt, magSfTri);
                }
            }
            return dVf;
        } else {
            if (mesh_.faceFlatness[faceI] > (1.0 - TSMALL)) {
                calcSubFace(faceI, normal, distance);
                const scalar alphaf = mag(subFaceArea_) / magSf;
                return (phi * dt * alphaf);
            } else {
                std::vector<point> fPtsTri(3);
                fPtsTri[0] = mesh_.Cf[faceI];
                scalar dVf = 0.0;
                for (label pi = 0; pi < nPoints; ++pi) {
                    fPtsTri[1] = fPts[pi];
                    fPtsTri[2] = fPts[(pi + 1) % nPoints];
                    const scalar magSfTri = mag(0.5 * ((fPtsTri[1] - fPtsTri[0]) ^ (fPtsTri[2] - fPtsTri[0])));
                    const scalar phiTri = phi * magSfTri / magSf;
                    calcSubFace(fPtsTri, normal, distance);
                    const scalar alphafTri = mag(subFaceArea_) / magSfTri;
                    dVf += (phiTri * dt * alphafTri);
                }
                return dVf;
            }
        }
    }

    // cutFace.C:392-506
    scalar timeIntegratedArea(const std::vector<point>& fPts, const vec& normal, scalar distance,
                              const std::vector<scalar>& pTimes, scalar Un0, scalar dt, scalar magSf)
    {
        scalar tIntArea = 0.0;
        const std::vector<label> order(sortedOrderLess(pTimes));
        const scalar firstTime = pTimes[order.front()];
        const scalar lastTime = pTimes[order.back()];

        if (lastTime <= 0.0) {
            tIntArea = magSf * dt * pos0(Un0);
            return tIntArea;
        }
        if (firstTime >= dt) {
            tIntArea = magSf * dt * (1.0 - pos0(Un0));
            return tIntArea;
        }

        std::vector<scalar> sortedTimes;
        scalar prevTime = 0.0;
        scalar subAreaOld = 0.0, subAreaNew = 0.0, subAreaMid = 0.0;

        if (firstTime > 0.0) {
            subAreaOld = magSf * (scalar(1) - pos0(Un0));
            tIntArea = subAreaOld * firstTime;
            sortedTimes.push_back(firstTime);
            prevTime = firstTime;
        } else {
            sortedTimes.push_back(0.0);
            prevTime = 0.0;
            calcSubFace(fPts, normal, distance);
            subAreaOld = mag(subFaceArea_);
        }

        const scalar smallTime = smax(TSMALL / mag(Un0), TSMALL);

        for (size_t ti = 0; ti < order.size(); ++ti) {
            const scalar timeI = pTimes[order[ti]];
            if (timeI > (prevTime + smallTime) && timeI < dt) {
                sortedTimes.push_back(timeI);
                prevTime = timeI;
            }
        }

        if (lastTime > dt) {
            sortedTimes.push_back(dt);
        } else {
            tIntArea += magSf * (dt - lastTime) * pos0(Un0);
        }

        for (int k = 0; k < int(sortedTimes.size()) - 1; k++) {
            const scalar tauOld = sortedTimes[k];
            const scalar tauNew = sortedTimes[k + 1];
            const scalar deltaTau = 0.5 * (tauNew - tauOld);

            calcSubFace(fPts, normal, distance - tauNew * Un0);
            subAreaNew = mag(subFaceArea_);

            calcSubFace(fPts, normal, distance - (tauOld + deltaTau) * Un0);
            subAreaMid = mag(subFaceArea_);

            tIntArea += (deltaTau / 3.0) * (subAreaOld + 4.0 * subAreaMid + subAreaNew);
            subAreaOld = subAreaNew;
        }
        return tIntArea;
    }

    const point& subFaceCentre() const { return subFaceCentre_; }
    const vec& subFaceArea() const { return subFaceArea_; }
    const std::vector<point>& subFacePoints() const { return subFacePoints_; }
    const std::vector<point>& interfacePoints() const { return interfacePoints_; }
};

// ---------------------------------------------------------------- cutCell ----
class cutCell {
    const Mesh& mesh_;
    cutFace cutFace_;
    label cellStatus_;
    std::vector<point> cutFaceCentres_;
    std::vector<vec> cutFaceAreas_;
    point subCellCentre_;
    scalar subCellVolume_;
    scalar VOF_;
    label cellI_;
    std::vector<std::vector<point>> interfaceEdges_;
    point interfaceCentre_;
    vec interfaceArea_;
    // splitWarpedFace == true
    std::vector<point> localPoints_;
    std::vector<std::vector<label>> localFaces_;

    // cutCell.C:37-100
    void calcInterfaceCentreAndArea()
    {
        point fCentre;
        label nEdgePoints = 0;
        for (const std::vector<point>& edgePoints : interfaceEdges_) {
            for (const point& p : edgePoints) {
                fCentre += p;
                nEdgePoints++;
            }
        }
        if (nEdgePoints > 0) fCentre /= scalar(nEdgePoints);

        vec sumN;
        scalar sumA = 0.0;
        vec sumAc;
        for (size_t ei = 0; ei < interfaceEdges_.size(); ++ei) {
            const std::vector<point>& edgePoints = interfaceEdges_[ei];
            const label nPoints = label(edgePoints.size());
            for (label pi = 0; pi < nPoints - 1; pi++) {
                const point& nextPoint = edgePoints[pi + 1];
                vec c(edgePoints[pi] + nextPoint + fCentre);
                vec n((nextPoint - edgePoints[pi]) ^ (fCentre - edgePoints[pi]));
                scalar a = mag(n);
                sumN += sign(n & sumN) * n;
                sumA += a;
                sumAc += a * c;
            }
        }
        if (sumA < ROOTVSMALL) {
            interfaceCentre_ = fCentre;
            interfaceArea_ = vec();
        } else {
            interfaceCentre_ = (1.0 / 3.0) * sumAc / sumA;
            interfaceArea_ = 0.5 * sumN;
        }
        // subCellCentre_ is still (0,0,0) here (clearStorage; SURVEY 8a' item 24)
        if ((interfaceArea_ & (interfaceCentre_ - subCellCentre_)) < 0.0) interfaceArea_ *= (-1.0);
    }

    // cutCell.C:103-137
    void calcSubCellCentreAndVolume()
    {
        subCellCentre_ = point();
        subCellVolume_ = 0.0;
        vec cEst;  // average(cutFaceCentres_) = sum/size
        for (const point& p : cutFaceCentres_) cEst += p;
        cEst /= scalar(cutFaceCentres_.size());
        for (size_t facei = 0; facei < cutFaceCentres_.size(); ++facei) {
            scalar pyr3Vol = smax(mag(cutFaceAreas_[facei] & (cutFaceCentres_[facei] - cEst)), VSMALL);
            vec pc(0.75 * cutFaceCentres_[facei] + 0.25 * cEst);
            subCellCentre_ += pyr3Vol * pc;
            subCellVolume_ += pyr3Vol;
        }
        subCellCentre_ /= subCellVolume_;
        subCellVolume_ /= 3.0;
    }

    // cutCell.C:140-236
    void getLocalPointFieldAndFaceList(bool splitWarpedFace)
    {
        localPoints_.clear();
        localFaces_.clear();
        // cell::labels(faces): unique point labels in order of first appearance
        std::vector<label> globalPointLabels;
        const label* c = mesh_.cells.row(cellI_);
        const label nc = mesh_.cells.size(cellI_);
        for (label fi = 0; fi < nc; ++fi) {
            const label f = c[fi];
            for (label k = 0; k < mesh_.faces.size(f); ++k) {
                const label pl = mesh_.faces.row(f)[k];
                if (std::find(globalPointLabels.begin(), globalPointLabels.end(), pl) == globalPointLabels.end())
                    globalPointLabels.push_back(pl);
            }
        }
        for (label pl : globalPointLabels) localPoints_.push_back(mesh_.points[pl]);
        auto findLocal = [&](label pl) {
            return label(std::find(globalPointLabels.begin(), globalPointLabels.end(), pl) - globalPointLabels.begin());
        };
        auto reverseFace = [](const std::vector<label>& f) {  // face::reverseFace: keeps vertex 0
            std::vector<label> r(f.size());
            r[0] = f[0];
            for (size_t i = 1; i < f.size(); ++i) r[i] = f[f.size() - i];
            return r;
        };
        for (label fi = 0; fi < nc; ++fi) {
            const label f = c[fi];
            const label* fa = mesh_.faces.row(f);
            const label fn = mesh_.faces.size(f);
            const bool own = (cellI_ == mesh_.owner[f]);
            if (!splitWarpedFace || mesh_.faceFlatness[f] > (1.0 - TSMALL)) {
                std::vector<label> localFa;
                for (label k = 0; k < fn; ++k) localFa.push_back(findLocal(fa[k]));
                localFaces_.push_back(own ? localFa : reverseFace(localFa));
            } else {
                localPoints_.push_back(mesh_.Cf[f]);
                for (label k = 0; k < fn; ++k) {
                    const label nextk = (k + 1) % fn;
                    std::vector<label> localFa(3);
                    localFa[0] = label(localPoints_.size()) - 1;
                    localFa[1] = findLocal(fa[k]);
                    localFa[2] = findLocal(fa[nextk]);
                    localFaces_.push_back(own ? localFa : reverseFace(localFa));
                }
            }
        }
    }

   public:
    bool collectSubCell_ = false;
    std::vector<point> subCellPoints_;
    std::vector<std::vector<label>> subCellFaces_;

    explicit cutCell(const Mesh& mesh)
        : mesh_(mesh), cutFace_(mesh), cellStatus_(-1), subCellVolume_(0), VOF_(0), cellI_(-1)
    {
        clearStorage();
    }

    void clearStorage()  // cutCell.H:291-306
    {
        cellI_ = -1;
        cellStatus_ = -1;
        cutFaceCentres_.clear();
        cutFaceAreas_.clear();
        interfaceEdges_.clear();
        interfaceCentre_ = point();
        interfaceArea_ = vec();
        subCellCentre_ = point();
        subCellVolume_ = -10;
        VOF_ = -10;
    }

    // cutCell.C:343-542 (sub-cell point/face lists, :382-393,:443-454, are only
    // consumed by surface export and are not restated)
    label calcSubCell(label cellI, const vec& normal, scalar distance, bool splitWarpedFace)
    {
        clearStorage();
        cellI_ = cellI;

        bool fullySubmerged = true;
        bool fullyEmpty = true;
        label nSubmergedFaces = 0;

        auto account = [&](label faceStatus) {
            if (faceStatus == 0) {
                cutFaceCentres_.push_back(cutFace_.subFaceCentre());
                cutFaceAreas_.push_back(cutFace_.subFaceArea());
                interfaceEdges_.push_back(cutFace_.interfacePoints());
                fullySubmerged = false;
                fullyEmpty = false;
            } else if (faceStatus == -1) {
                cutFaceCentres_.push_back(cutFace_.subFaceCentre());
                cutFaceAreas_.push_back(cutFace_.subFaceArea());
                fullyEmpty = false;
                nSubmergedFaces++;
            } else {
                fullySubmerged = false;
            }
        };

        if (splitWarpedFace) {
            for (size_t i = 0; i < localFaces_.size(); ++i) {
                std::vector<point> fPts(localFaces_[i].size());
                for (size_t k = 0; k < fPts.size(); ++k) fPts[k] = localPoints_[localFaces_[i][k]];
                account(cutFace_.calcSubFace(fPts, normal, distance));
            }
        } else {
            const label* c = mesh_.cells.row(cellI);
            subCellPoints_.clear();
            subCellFaces_.clear();
            for (label fi = 0; fi < mesh_.cells.size(cellI); ++fi) {
                const label st = cutFace_.calcSubFace(c[fi], normal, distance);
                account(st);
                if (collectSubCell_ && st <= 0) {  // cutCell.C:443-454 (cut face), :464-475 (fully submerged face)
                    const std::vector<point> fp = (st == 0) ? cutFace_.subFacePoints() : cutFace_.facePoints(c[fi]);
                    std::vector<label> f(fp.size());
                    for (size_t k = 0; k < fp.size(); ++k) f[k] = label(subCellPoints_.size() + k);
                    subCellFaces_.push_back(f);
                    subCellPoints_.insert(subCellPoints_.end(), fp.begin(), fp.end());
                }
            }
        }

        if (!fullySubmerged && !fullyEmpty) {
            cellStatus_ = 0;
            calcInterfaceCentreAndArea();
            if (mag(interfaceArea_) < TSMALL) {
                if (nSubmergedFaces == 0) {
                    cellStatus_ = 1;
                    subCellCentre_ = point();
                    subCellVolume_ = 0.0;
                    VOF_ = 0.0;
                    return cellStatus_;
                } else {
                    cellStatus_ = -1;
                    subCellCentre_ = mesh_.C[cellI];
                    subCellVolume_ = mesh_.V[cellI];
                    VOF_ = 1.0;
                    return cellStatus_;
                }
            }
            cutFaceCentres_.push_back(interfaceCentre_);
            cutFaceAreas_.push_back(interfaceArea_);
            calcSubCellCentreAndVolume();
            VOF_ = subCellVolume_ / mesh_.V[cellI];
        } else if (fullyEmpty) {
            cellStatus_ = 1;
            subCellCentre_ = point();
            subCellVolume_ = 0.0;
            VOF_ = 0.0;
        } else if (fullySubmerged) {
            cellStatus_ = -1;
            subCellCentre_ = mesh_.C[cellI];
            subCellVolume_ = mesh_.V[cellI];
            VOF_ = 1.0;
        }
        return cellStatus_;
    }

    // cutCell.C:611-799.  Writes D/C/S only where the reference does.
    label findSignedDistance(label cellI, scalar alphaI, const vec& normalI, bool splitWarpedFace, scalar& interfaceD,
                             vec& interfaceC, vec& interfaceS)
    {
        cellI_ = cellI;
        if (mag(normalI) < TSMALL) return label(sign(0.5 - alphaI));

        std::vector<scalar> vertexDistances;
        if (splitWarpedFace) {
            getLocalPointFieldAndFaceList(true);
            vertexDistances.resize(localPoints_.size());
            for (size_t pointi = 0; pointi < localPoints_.size(); ++pointi)
                vertexDistances[pointi] = -(normalI & localPoints_[pointi]);
        } else {
            const label* pLabels = mesh_.cellPoints.row(cellI);
            vertexDistances.resize(mesh_.cellPoints.size(cellI));
            for (size_t pointi = 0; pointi < vertexDistances.size(); ++pointi)
                vertexDistances[pointi] = -(normalI & mesh_.points[pLabels[pointi]]);
        }
        const std::vector<label> vertexDistanceOrder(sortedOrderGreater(vertexDistances));

        scalar lowDistance = vertexDistances[vertexDistanceOrder.front()];
        scalar upDistance = vertexDistances[vertexDistanceOrder.back()];
        label lowLabel = 0;
        label upLabel = label(vertexDistances.size()) - 1;
        scalar lowAlpha = 0.0;
        scalar upAlpha = 1.0;
        scalar midDistance, midLabel, midAlpha;  // midLabel IS a scalar in the reference (:689)

        while ((upLabel - lowLabel) > 1) {
            midLabel = std::round(0.5 * (upLabel + lowLabel));
            midDistance = vertexDistances[vertexDistanceOrder[label(midLabel)]];
            calcSubCell(cellI, normalI, midDistance, splitWarpedFace);
            midAlpha = VOF_;
            if (mag(midAlpha - alphaI) < TSMALL) {
                interfaceD = midDistance;
                interfaceC = interfaceCentre_;
                interfaceS = interfaceArea_;
                return cellStatus_;
            }
            if (midAlpha > alphaI) {
                upLabel = label(midLabel);
                upDistance = midDistance;
                upAlpha = midAlpha;
            } else {
                lowLabel = label(midLabel);
                lowDistance = midDistance;
                lowAlpha = midAlpha;
            }
        }

        if (mag(lowDistance - upDistance) < TSMALL) {
            const scalar midD = 0.5 * (lowDistance + upDistance);
            calcSubCell(cellI, normalI, midD, splitWarpedFace);
            interfaceD = midD;
            interfaceC = interfaceCentre_;
            interfaceS = interfaceArea_;
            return cellStatus_;
        }

        const scalar alphaPrismatoid = upAlpha - lowAlpha;
        const scalar deltaDistance = (upDistance - lowDistance) / 3.0;

        const scalar distanceOneThird = lowDistance + deltaDistance;
        calcSubCell(cellI, normalI, distanceOneThird, splitWarpedFace);
        const scalar alphaOneThird = VOF_ - lowAlpha;

        const scalar distanceTwoThirds = lowDistance + 2.0 * deltaDistance;
        calcSubCell(cellI, normalI, distanceTwoThirds, splitWarpedFace);
        const scalar alphaTwoThirds = VOF_ - lowAlpha;

        scalar a, b, c, d;
        a = 13.5 * alphaOneThird - 13.5 * alphaTwoThirds + 4.5 * alphaPrismatoid;
        b = -22.5 * alphaOneThird + 18.0 * alphaTwoThirds - 4.5 * alphaPrismatoid;
        c = 9.0 * alphaOneThird - 4.5 * alphaTwoThirds + 1.0 * alphaPrismatoid;
        d = lowAlpha - alphaI;

        scalar lambda = 0.5;
        for (label iter = 0; iter < 100; iter++) {
            const scalar func = a * pow3(lambda) + b * sqr(lambda) + c * lambda + d;
            const scalar funcPrime = 3.0 * a * sqr(lambda) + 2.0 * b * lambda + c;
            const scalar lambdaNew = lambda - (func / funcPrime);
            if (mag(lambdaNew - lambda) < TSMALL) break;
            lambda = lambdaNew;
        }

        const scalar distance0 = lowDistance - lambda * (lowDistance - upDistance);
        calcSubCell(cellI, normalI, distance0, splitWarpedFace);
        interfaceD = distance0;
        interfaceC = interfaceCentre_;
        interfaceS = interfaceArea_;
        return cellStatus_;
    }

    // mapAlphaField/interface()/subCellFaces() call calcSubCell with
    // splitWarpedFace=false (reconstruction.C:768,805,856)
    scalar volumeOfFluid() const { return VOF_; }
    scalar subCellVolume() const { return subCellVolume_; }
    const point& subCellCentre() const { return subCellCentre_; }
    const point& interfaceCentre() const { return interfaceCentre_; }

    // cutCell::interfacePoints (cutCell.C:545-608): the interface polygon of the last calcSubCell, its points sorted by
    // angle about the interface centre in the plane of the interface; points closer than 1e-8 rad are merged.
    // sortedOrder is a stable ascending sort (OF, recalled).
    void interfacePolygon(std::vector<point>& out) const
    {
        out.clear();
        if (cellStatus_ != 0 || interfaceEdges_.empty()) return;
        const vec zhat = interfaceArea_ / mag(interfaceArea_);
        vec xhat = interfaceEdges_[0][0] - interfaceCentre_;
        xhat = (xhat - (xhat & zhat) * zhat);
        xhat /= mag(xhat);      // Vector::normalise(): divides when mag >= ROOTVSMALL (OF, recalled)
        vec yhat = zhat ^ xhat;
        yhat /= mag(yhat);
        std::vector<point> pts;
        std::vector<scalar> ang;
        for (const std::vector<point>& edgePoints : interfaceEdges_)
            for (const point& p : edgePoints) {
                pts.push_back(p);
                ang.push_back(std::atan2((p - interfaceCentre_) & yhat, (p - interfaceCentre_) & xhat));
            }
        std::vector<label> order(pts.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = label(i);
        std::stable_sort(order.begin(), order.end(), [&](label a, label b) { return ang[a] < ang[b]; });
        out.push_back(pts[order[0]]);
        for (size_t pi = 1; pi < order.size(); ++pi)
            if (std::fabs(ang[order[pi]] - ang[order[pi - 1]]) > 1e-8) out.push_back(pts[order[pi]]);
    }
    const vec& interfaceArea() const { return interfaceArea_; }
    label cellStatus() const { return cellStatus_; }

    // cutCell::updateSubCellPointsandFaces (cutCell.C:239-290) for the last calcSubCell (status 0, collectSubCell(true)):
    // the submerged sub-faces plus the interface polygon, duplicate points merged (inplaceMergePoints, 10*SMALL; kept in
    // order of first occurrence -- OF's internal order is not recalled), every face oriented away from the sub-cell centre.
    void collectSubCell(bool on) { collectSubCell_ = on; }
    void subCellPointsAndFaces(std::vector<point>& pts, std::vector<std::vector<label>>& faces) const
    {
        std::vector<point> all(subCellPoints_);
        faces = subCellFaces_;
        std::vector<point> poly;
        interfacePolygon(poly);
        if (!poly.empty()) {
            std::vector<label> f(poly.size());
            for (size_t k = 0; k < poly.size(); ++k) f[k] = label(all.size() + k);
            faces.push_back(f);
            all.insert(all.end(), poly.begin(), poly.end());
        }
        std::vector<label> toUnique(all.size());
        pts.clear();
        for (size_t i = 0; i < all.size(); ++i) {
            label found = -1;
            for (size_t j = 0; j < pts.size() && found < 0; ++j)
                if (mag(all[i] - pts[j]) <= 10.0 * SMALL) found = label(j);
            if (found < 0) {
                found = label(pts.size());
                pts.push_back(all[i]);
            }
            toUnique[i] = found;
        }
        for (std::vector<label>& f : faces) {
            for (label& v : f) v = toUnique[v];
            std::vector<point> fp(f.size());
            for (size_t k = 0; k < f.size(); ++k) fp[k] = pts[f[k]];
            const point fc = faceCentreOF(fp);
            const vec fn = faceAreaNormalOF(fp);
            if (((fc - subCellCentre_) & fn) < 0.0) std::reverse(f.begin() + 1, f.end());  // face::reverseFace keeps vertex 0
        }
    }
    cutFace& faceCutter() { return cutFace_; }
};

}  // namespace ora
