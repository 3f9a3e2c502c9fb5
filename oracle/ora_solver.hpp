// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Restatement of the reference's L2 classes on plain arrays:
//   reconstruction <- src/SimPLIC/reconstruction/reconstruction.C
//   advection      <- src/SimPLIC/advection/advection.C, advectionTemplates.C
// plus the OpenFOAM services they call (OF v2312, not vendored -- "OF,
// recalled"): leastSquareGrad/multiDimPolyFitter + LUsolve,
// zoneCPCStencil membership, volPointInterpolation, interpolationCellPoint,
// upwind::flux, fvc::surfaceIntegrate, zeroGradient/fixedValue/inletOutlet.
//
// PARITY STATUS: the reference tree stores no SimPLIC output, and
// OpenFOAM is absent, so per-cell results are "parity unpinned" against the
// real binary; what IS pinned (tests/test_oracle_golden.py) are the reference's
// golden exact fields, exactInitialVol and the metric definitions.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "ora_cut.hpp"

namespace ora {

static const scalar aTol = 100.0 * SMALL;  // advectionTemplates.C:125,228

// ---- Foam::LUDecompose / LUBacksubstitute (OF scalarMatrices.C, recalled) ----
// Crout LU with implicit-scaling partial pivoting, in place, n <= 4.
inline void LUsolve(scalar A[4][4], scalar b[4], int m)
{
    int pivot[4];
    scalar vv[4];
    for (int i = 0; i < m; ++i) {
        scalar largestCoeff = 0.0, temp;
        for (int j = 0; j < m; ++j)
            if ((temp = mag(A[i][j])) > largestCoeff) largestCoeff = temp;
        if (largestCoeff == 0.0) largestCoeff = SMALL;  // OF aborts on a singular matrix; keep finite
        vv[i] = 1.0 / largestCoeff;
    }
    for (int j = 0; j < m; ++j) {
        for (int i = 0; i < j; ++i) {
            scalar sum = A[i][j];
            for (int k = 0; k < i; ++k) sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
        }
        int iMax = 0;
        scalar largestCoeff = 0.0;
        for (int i = j; i < m; ++i) {
            scalar sum = A[i][j];
            for (int k = 0; k < j; ++k) sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
            scalar temp;
            if ((temp = vv[i] * mag(sum)) >= largestCoeff) {
                largestCoeff = temp;
                iMax = i;
            }
        }
        pivot[j] = iMax;
        if (j != iMax) {
            for (int k = 0; k < m; ++k) std::swap(A[j][k], A[iMax][k]);
            vv[iMax] = vv[j];
        }
        if (A[j][j] == 0.0) A[j][j] = SMALL;
        if (j != m - 1) {
            scalar rDiag = 1.0 / A[j][j];
            for (int i = j + 1; i < m; ++i) A[i][j] *= rDiag;
        }
    }
    int ii = 0;
    for (int i = 0; i < m; ++i) {
        int ip = pivot[i];
        scalar sum = b[ip];
        b[ip] = b[i];
        if (ii != 0) {
            for (int j = ii - 1; j < i; ++j) sum -= A[i][j] * b[j];
        } else if (sum != 0.0) {
            ii = i + 1;
        }
        b[i] = sum;
    }
    for (int i = m - 1; i >= 0; --i) {
        scalar sum = b[i];
        for (int j = i + 1; j < m; ++j) sum -= A[i][j] * b[j];
        b[i] = sum / A[i][i];
    }
}

struct Solver {
    Mesh mesh;
    svof_params prm;
    std::string err;

    // fields
    std::vector<scalar> alpha, alphaOld, alphaB, phi;
    std::vector<vec> U, Ub;
    bool haveAlpha = false, havePhi = false, haveU = false;
    // reconstruction state
    std::vector<vec> interfaceN, interfaceC, interfaceS;
    std::vector<scalar> interfaceD;
    std::vector<label> mixedCells, cellStatus;
    std::vector<scalar> Un0;  // per mixed cell, for tests
    // advection state
    std::vector<scalar> dVf, alphaPhi;
    // info
    scalar minBefore = 0, maxM1Before = 0, minAfter = 0, maxM1After = 0;
    label nSweeps = 0;
    double reconstructionTime = 0, advectionTime = 0;
    int geomD[3] = {1, 1, 1};

    cutCell* cutCell_ = nullptr;
    cutFace* cutFace_ = nullptr;

    ~Solver()
    {
        delete cutCell_;
        delete cutFace_;
    }

    void init(const svof_mesh& m, const svof_params& p)
    {
        prm = p;
        mesh.build(m);
        for (const svof_patch& pt : mesh.patches)
            if (pt.kind == SVOF_PATCH_PROCESSOR && pt.size > 0)
                throw std::invalid_argument("the CPU oracle is single-domain: processor patches are not supported");
        const label nC = mesh.nCells, nF = mesh.nFaces;
        alpha.assign(nC, 0);
        alphaOld.assign(nC, 0);
        alphaB.assign(mesh.nBoundaryFaces(), 0);
        phi.assign(nF, 0);
        U.assign(nC, vec());
        Ub.assign(mesh.nBoundaryFaces(), vec());
        interfaceN.assign(nC, vec());
        interfaceC.assign(nC, vec());
        interfaceS.assign(nC, vec());
        interfaceD.assign(nC, 0);
        dVf.assign(nF, 0);
        alphaPhi.assign(nF, 0);
        cutCell_ = new cutCell(mesh);
        cutFace_ = new cutFace(mesh);
        // polyMesh::geometricD(): directions normal to empty patches are -1 (OF, recalled)
        vec emptyDirVec;
        for (const svof_patch& pt : mesh.patches) {
            if (pt.kind != SVOF_PATCH_EMPTY) continue;
            for (label k = 0; k < pt.size; ++k) {
                const label f = pt.start + k;
                if (mesh.magSf[f] > 0) {
                    vec nf = mesh.Sf[f] / mesh.magSf[f];
                    emptyDirVec += vec(mag(nf.x), mag(nf.y), mag(nf.z));
                }
            }
        }
        if (mag(emptyDirVec) > 0) {
            emptyDirVec /= mag(emptyDirVec);
            geomD[0] = emptyDirVec.x > 1e-6 ? -1 : 1;
            geomD[1] = emptyDirVec.y > 1e-6 ? -1 : 1;
            geomD[2] = emptyDirVec.z > 1e-6 ? -1 : 1;
        }
    }

    // volScalarField::correctBoundaryConditions for the supported patch types
    void correctAlphaBCs()
    {
        const label nIF = mesh.nInternalFaces;
        for (const svof_patch& pt : mesh.patches) {
            for (label k = 0; k < pt.size; ++k) {
                const label f = pt.start + k, bf = f - nIF;
                if (pt.kind != SVOF_PATCH_GENERIC) {
                    alphaB[bf] = 0;
                    continue;
                }
                const scalar internal = alpha[mesh.owner[f]];
                switch (pt.alpha_bc) {
                    case SVOF_BC_FIXED_VALUE: alphaB[bf] = pt.alpha_value; break;
                    case SVOF_BC_INLET_OUTLET: {
                        // inletOutletFvPatchField: valueFraction = 1 - pos0(phi)
                        const scalar vf = 1.0 - pos0(phi[f]);
                        alphaB[bf] = vf * pt.alpha_value + (1.0 - vf) * internal;
                        break;
                    }
                    default: alphaB[bf] = internal;
                }
            }
        }
    }

    bool isAMixedCell(label c) const  // reconstruction.H:281-288
    {
        return (prm.mixed_cell_tol < alpha[c]) && (alpha[c] < 1.0 - prm.mixed_cell_tol);
    }

    // overset: cellCellStencil cell types (empty = not an overset mesh); CALCULATED = 0
    std::vector<label> cellTypes;

    // ---------------------------------------------------------------- A1 ----
    void initialize()  // reconstruction.C:634-677
    {
        mixedCells.clear();
        cellStatus.clear();
        std::fill(interfaceN.begin(), interfaceN.end(), vec());
        std::fill(interfaceC.begin(), interfaceC.end(), vec());
        std::fill(interfaceS.begin(), interfaceS.end(), vec());
        std::fill(interfaceD.begin(), interfaceD.end(), 0.0);
        const bool overset = !cellTypes.empty();   // isA<dynamicOversetFvMesh>(mesh_)  (:649-662)
        for (label c = 0; c < mesh.nCells; ++c) {
            if (isAMixedCell(c) && (!overset || cellTypes[c] == 0)) {
                mixedCells.push_back(c);
                cellStatus.push_back(-100);
            }
        }
    }

    // zoneCPCStencil membership (OF, recalled): the cell itself first, then every
    // cell sharing a vertex, then the valid (non-empty, non-coupled) boundary
    // faces at its vertices as pseudo-cells nCells+bFace.  OF iterates a hash
    // set, i.e. its order is implementation-defined; the oracle (and the CUDA
    // path) pin ascending label order (SURVEY.md 8e "Determinism").
    void cpcStencil(label celli, std::vector<label>& st) const
    {
        st.clear();
        std::vector<label> cs, bs;
        for (label k = 0; k < mesh.cellPoints.size(celli); ++k) {
            const label p = mesh.cellPoints.row(celli)[k];
            for (label j = 0; j < mesh.pointCells.size(p); ++j) {
                const label c = mesh.pointCells.row(p)[j];
                if (c != celli) cs.push_back(c);
            }
            for (label j = 0; j < mesh.pointBFaces.size(p); ++j) {
                const label bf = mesh.pointBFaces.row(p)[j];
                if (mesh.isPatchFace[bf]) bs.push_back(bf);
            }
        }
        std::sort(cs.begin(), cs.end());
        cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
        std::sort(bs.begin(), bs.end());
        bs.erase(std::unique(bs.begin(), bs.end()), bs.end());
        st.push_back(celli);
        st.insert(st.end(), cs.begin(), cs.end());
        for (label bf : bs) st.push_back(mesh.nCells + bf);
    }

    // ---------------------------------------------------------------- A2 ----
    // reconstruction.C:74-82: interfaceN = -fvc::grad(alpha1) with the case's gradScheme, which every shipped case sets
    // to `Gauss linear` (e.g. tutorials/test/plicVofAdvectionFoam/system/fvSchemes).  OF, recalled:
    //   surfaceInterpolation::makeWeights   w = |Sf.(C_N - Cf)| / (|Sf.(Cf - C_P)| + |Sf.(C_N - Cf)|)
    //   linear interpolate                  a_f = w*(a_P - a_N) + a_N ; boundary faces take the patch value
    //   GaussGrad::calcGrad                 faces ascending: grad[P] += Sf*a_f, grad[N] -= Sf*a_f; patches; then /V
    // With `Gauss pointLinear` (tutorials/test/plicVofOrientationFoam/NAG/system/fvSchemes:35; prm.alpha_grad_scheme 1) the face
    // value is linear + pointLinear::correction.
    // Only the mixed cells are evaluated: the reference fills every cell, but nothing on the path reads the others.
    // surfaceInterpolation::weights() of internal face f (OF, recalled)
    scalar linearWeight(label f) const
    {
        const label P = mesh.owner[f], N = mesh.neighbour[f];
        const scalar SfdOwn = std::fabs(mesh.Sf[f] & (mesh.Cf[f] - mesh.C[P]));
        const scalar SfdNei = std::fabs(mesh.Sf[f] & (mesh.C[N] - mesh.Cf[f]));
        return (std::fabs(SfdOwn + SfdNei) > ROOTVSMALL) ? SfdNei / (SfdOwn + SfdNei) : 0.5;
    }

    // volPointInterpolation::interpolate(alpha1) evaluated lazily at one point (as pointU below)
    scalar pointAlpha(label p) const
    {
        const point& pt = mesh.points[p];
        if (!mesh.isPatchPoint[p]) {
            const label n = mesh.pointCells.size(p);
            const label* pc = mesh.pointCells.row(p);
            scalar sumW = 0.0;
            for (label k = 0; k < n; ++k) sumW += 1.0 / mag(pt - mesh.C[pc[k]]);
            scalar val = 0.0;
            for (label k = 0; k < n; ++k) {
                const scalar pw = (1.0 / mag(pt - mesh.C[pc[k]])) / sumW;
                val += pw * alpha[pc[k]];
            }
            return val;
        }
        const label n = mesh.pointBFaces.size(p);
        const label* pf = mesh.pointBFaces.row(p);
        scalar sumW = 0.0;
        for (label k = 0; k < n; ++k)
            if (mesh.isPatchFace[pf[k]]) sumW += 1.0 / mag(pt - mesh.Cf[mesh.nInternalFaces + pf[k]]);
        scalar val = 0.0;
        for (label k = 0; k < n; ++k) {
            if (!mesh.isPatchFace[pf[k]]) continue;
            const scalar pw = (1.0 / mag(pt - mesh.Cf[mesh.nInternalFaces + pf[k]])) / sumW;
            val += pw * alphaB[pf[k]];
        }
        return val;
    }

    // pointLinear<scalar>::correction on internal face f (OF, recalled: pointLinear.C): the face value is re-built from the
    // point-interpolated field over the triangles (pi, f[k], f[k-1]) about pi = a C_P + (1 - a) C_N; `lin` is
    // linearInterpolate(vf)[f].  As published, a is mesh.weights() indexed by the OWNER CELL label (not by the face); the
    // quirk is kept (a = 0.5 where that index is not an internal face).
    scalar pointLinearCorrection(label f, scalar lin) const
    {
        const label P = mesh.owner[f], N = mesh.neighbour[f];
        const scalar a = (P < mesh.nInternalFaces) ? linearWeight(P) : 0.5;
        const point pi = a * mesh.C[P] + (1.0 - a) * mesh.C[N];
        const label nv = mesh.faces.size(f);
        const label* fp = mesh.faces.row(f);
        auto triMag = [&](const point& b, const point& c) { return mag(0.5 * ((b - pi) ^ (c - pi))); };
        scalar at = triMag(mesh.points[fp[0]], mesh.points[fp[nv - 1]]);
        scalar sumAt = at;
        scalar sumPsip = at * (1.0 / 3.0) * (lin + pointAlpha(fp[0]) + pointAlpha(fp[nv - 1]));
        for (label k = 1; k < nv; ++k) {
            at = triMag(mesh.points[fp[k]], mesh.points[fp[k - 1]]);
            sumAt += at;
            sumPsip += at * (1.0 / 3.0) * (lin + pointAlpha(fp[k]) + pointAlpha(fp[k - 1]));
        }
        return sumPsip / sumAt - lin;
    }

    void calcInterfaceNFromRegAlphaGrad()
    {
        std::vector<label> fs;
        const bool pointLinear = prm.alpha_grad_scheme == 1;
        for (size_t i = 0; i < mixedCells.size(); ++i) {
            const label celli = mixedCells[i];
            fs.assign(mesh.cells.row(celli), mesh.cells.row(celli) + mesh.cells.size(celli));
            std::sort(fs.begin(), fs.end());
            vec g;
            for (label f : fs) {
                if (f < mesh.nInternalFaces) {
                    const label P = mesh.owner[f], N = mesh.neighbour[f];
                    const scalar w = linearWeight(f);
                    scalar af = w * (alpha[P] - alpha[N]) + alpha[N];
                    if (pointLinear) af += pointLinearCorrection(f, af);   // surfaceInterpolationScheme::interpolate: tsf += correction
                    const vec Sfssf = mesh.Sf[f] * af;
                    if (celli == P) g += Sfssf;
                    else g -= Sfssf;
                } else {
                    const label bf = f - mesh.nInternalFaces;
                    if (!mesh.isPatchFace[bf]) continue;  // empty patches have no faces in fvMesh
                    g += mesh.Sf[f] * alphaB[bf];          // (the correction is zero on uncoupled patches)
                }
            }
            g /= mesh.V[celli];
            interfaceN[celli] = -g;
            interfaceN[celli] /= (mag(interfaceN[celli]) + SMALL);
        }
    }

    // leastSquareGrad<scalar>("polyDegree1", geometricD).grad(positions - C[celli], values) over the stencil st
    // (multiDimPolyFitter::fitData: resetMatrix, fillMatrix per sample, LUsolve); cell values / boundary-face values
    vec lsGrad(label celli, const std::vector<label>& st, const std::vector<scalar>& cellVal, const std::vector<scalar>& bVal) const
    {
        int nDims = 0;
        for (int d = 0; d < 3; ++d) nDims += (geomD[d] == 1);
        const int nTerms = 1 + nDims;
        scalar A[4][4] = {{0}}, src[4] = {0};
        for (label g : st) {
            vec pos;
            scalar val;
            if (g < mesh.nCells) {
                pos = mesh.C[g];
                val = cellVal[g];
            } else {
                pos = mesh.Cf[mesh.nInternalFaces + (g - mesh.nCells)];
                val = bVal[g - mesh.nCells];
            }
            pos -= mesh.C[celli];  // cellCentre -= mesh_.C()[celli]  (:133)
            const scalar comp[3] = {pos.x, pos.y, pos.z};
            scalar terms[4];
            terms[0] = 1;  // polyDegree1::termValues
            int dimCounter = 0;
            for (int d = 0; d < 3; ++d)
                if (geomD[d] == 1) terms[++dimCounter] = comp[d];
            for (int r = 0; r < nTerms; ++r) {
                src[r] += terms[r] * val;
                for (int q = 0; q < nTerms; ++q) A[r][q] += terms[r] * terms[q];
            }
        }
        LUsolve(A, src, nTerms);
        scalar g3[3] = {0, 0, 0};
        int dimCounter = 0;
        for (int d = 0; d < 3; ++d)
            if (geomD[d] == 1) g3[d] = src[++dimCounter];
        return vec(g3[0], g3[1], g3[2]);
    }

    void calcInterfaceNFromIsoAlphaGrad()  // reconstruction.C:85-141
    {
        std::vector<label> st;
        for (size_t i = 0; i < mixedCells.size(); ++i) {
            const label celli = mixedCells[i];
            cpcStencil(celli, st);
            interfaceN[celli] = -lsGrad(celli, st, alpha, alphaB);  // :135
        }
        for (label c = 0; c < mesh.nCells; ++c) interfaceN[c] /= (mag(interfaceN[c]) + SMALL);  // :138
    }

    // Vector::normalise(tol) (OF, recalled): s = mag; s < tol ? Zero : v / s, in place
    static vec& normalise(vec& v, scalar tol)
    {
        const scalar s = mag(v);
        if (s < tol) v = vec(); else v /= s;
        return v;
    }

    // reconstructedDistanceFunction::constructRDF (OF v2312 src/transportModels/geometricVoF/reconstructedDistanceFunction,
    // recalled; called at reconstruction.C:257-264 with centre = interfaceC, normal = -interfaceN, updateStencil = false):
    // interface cells get the signed distance of their own plane, the other cells of the zone a weighted average of the
    // distances to the planes of their stencil cells, weight = cos^2 of the angle between the connection and the normal;
    // calculated patches get the same average at their face centres.
    void constructRDF(const std::vector<char>& nextToInterface, std::vector<scalar>& RDF, std::vector<scalar>& RDFb) const
    {
        std::vector<label> st;
        auto average = [&](const vec& p, const std::vector<label>& stc, scalar& averageDist, scalar& avgWeight) {
            averageDist = 0;
            avgWeight = 0;
            for (label g : stc) {
                if (g >= mesh.nCells) continue;  // boundary values of interfaceN are zero (calculated patches of a zero field)
                vec n = interfaceN[g];           // -(-interfaceN)
                if (mag(n) != 0) {
                    n /= mag(n);
                    const vec c = interfaceC[g];
                    vec distanceToIntSeg = c - p;
                    const scalar distToSurf = distanceToIntSeg & n;
                    scalar weight = 0;
                    if (mag(distanceToIntSeg) != 0) {
                        distanceToIntSeg /= mag(distanceToIntSeg);
                        const scalar m_ = std::fabs(distanceToIntSeg & n);
                        weight = m_ * m_;
                    } else {
                        weight = 1;
                    }
                    averageDist += distToSurf * weight;
                    avgWeight += weight;
                }
            }
        };
        for (label celli = 0; celli < mesh.nCells; ++celli) {
            if (!nextToInterface[celli]) {
                RDF[celli] = 0;
                continue;
            }
            if (mag(interfaceN[celli]) != 0) {  // interface cell
                const vec n = interfaceN[celli] / mag(interfaceN[celli]);
                RDF[celli] = (interfaceC[celli] - mesh.C[celli]) & n;
            } else {
                cpcStencil(celli, st);
                scalar averageDist, avgWeight;
                average(mesh.C[celli], st, averageDist, avgWeight);
                if (avgWeight != 0) RDF[celli] = averageDist / avgWeight;
            }
        }
        for (label bf = 0; bf < mesh.nBoundaryFaces(); ++bf) {
            if (!mesh.isPatchFace[bf]) continue;  // calculated patches only (empty / coupled keep their constraint type)
            const label pCellI = mesh.owner[mesh.nInternalFaces + bf];
            if (!nextToInterface[pCellI]) {
                RDFb[bf] = 0;
                continue;
            }
            cpcStencil(pCellI, st);
            scalar averageDist, avgWeight;
            average(mesh.Cf[mesh.nInternalFaces + bf], st, averageDist, avgWeight);
            RDFb[bf] = (avgWeight != 0) ? averageDist / avgWeight : 0;
        }
    }

    label isoRDFIterationsDone = 0;   // for tests: iterations the last calcInterfaceNFromIsoRDF executed
    void calcInterfaceNFromIsoRDF()  // reconstruction.C:196-405
    {
        const scalar TSMALL = 10.0 * SMALL;
        const size_t nMixed = mixedCells.size();
        std::vector<vec> interfaceNormal(nMixed);
        std::vector<char> isMixedCell(mesh.nCells, 0), nextToInterface(mesh.nCells, 0);
        for (label c : mixedCells) isMixedCell[c] = 1;
        // reconstructedDistanceFunction::markCellsNearSurf(isMixedCell, 1) (OF, recalled): the interface cells and
        // every cell sharing a vertex with one
        for (label c : mixedCells) {
            nextToInterface[c] = 1;
            for (label k = 0; k < mesh.cellPoints.size(c); ++k) {
                const label p = mesh.cellPoints.row(c)[k];
                for (label j = 0; j < mesh.pointCells.size(p); ++j) nextToInterface[mesh.pointCells.row(p)[j]] = 1;
            }
        }
        std::vector<std::vector<label>> stencil(nMixed);
        for (size_t i = 0; i < nMixed; ++i) cpcStencil(mixedCells[i], stencil[i]);
        for (size_t i = 0; i < nMixed; ++i) interfaceNormal[i] = lsGrad(mixedCells[i], stencil[i], alpha, alphaB);  // :219
        // interfaceC_, interfaceN_ = Zero (:223-224): already zero after initialize()
        std::vector<char> tooCoarse(mesh.nCells, 0);
        std::vector<scalar> RDF(mesh.nCells, 0.0), RDFb(mesh.nBoundaryFaces(), 0.0);   // a fresh RDF object per call (:204)
        std::vector<scalar> normalResNormalResidual(nMixed), normalResAvgAngle(nMixed);
        isoRDFIterationsDone = 0;
        for (label iter = 0; iter < prm.rdf_iterations; ++iter) {
            ++isoRDFIterationsDone;
            for (size_t i = 0; i < nMixed; ++i) {
                const label celli = mixedCells[i];
                interfaceN[celli] = -normalise(interfaceNormal[i], SMALL);
                if (tooCoarse[celli]) continue;
                cellStatus[i] = cutCell_->findSignedDistance(celli, alpha[celli], interfaceN[celli], prm.split_warped_face != 0,
                                                             interfaceD[celli], interfaceC[celli], interfaceS[celli]);
            }
            constructRDF(nextToInterface, RDF, RDFb);   // updateContactAngle (:266): no contact-angle patches on this path
            for (size_t i = 0; i < nMixed; ++i) interfaceNormal[i] = lsGrad(mixedCells[i], stencil[i], RDF, RDFb);  // :268
            for (size_t i = 0; i < nMixed; ++i) {
                const label celli = mixedCells[i];
                if (mag(interfaceN[celli]) < TSMALL || mag(normalise(interfaceNormal[i], SMALL)) < TSMALL) {
                    normalResNormalResidual[i] = 0.0;   // (:283-292 -- no `continue` in the reference: overwritten below)
                    normalResAvgAngle[i] = 0.0;
                }
                scalar avgDiffNormal = 0, maxDiffNormal = GREAT, weight = 0;
                const vec cellNormal = interfaceN[celli];
                for (size_t j = 0; j < stencil[i].size(); ++j) {
                    const label g = stencil[i][j];
                    const vec normal = (g < mesh.nCells) ? interfaceN[g] : vec();
                    if (mag(normal) >= TSMALL && j != 0) {
                        const vec n = normal / mag(normal);
                        const scalar cosAngle = std::max(std::min((cellNormal & n), 1.0), -1.0);
                        avgDiffNormal += std::acos(cosAngle) * mag(normal);
                        weight += mag(normal);
                        if (cosAngle < maxDiffNormal) maxDiffNormal = cosAngle;
                    }
                }
                if (weight != 0) avgDiffNormal /= weight; else avgDiffNormal = 0;
                const vec newCellNormal = -normalise(interfaceNormal[i], SMALL);
                normalResNormalResidual[i] = 1.0 - (cellNormal & newCellNormal);
                normalResAvgAngle[i] = avgDiffNormal;
            }
            label resCounter = 0;
            scalar avgRes = 0, avgNormRes = 0;
            for (size_t i = 0; i < nMixed; ++i) {
                const scalar normalRes = normalResNormalResidual[i], avgA = normalResAvgAngle[i];
                if (avgA > 0.26 && iter > 0) {  // 15 deg
                    tooCoarse[mixedCells[i]] = 1;
                } else {
                    avgRes += normalRes;
                    scalar normRes = 0;
                    const scalar discreteError = 0.01 * (avgA * avgA);
                    if (discreteError != 0) normRes = normalRes / std::max(discreteError, prm.rdf_tol);
                    else normRes = normalRes / prm.rdf_tol;
                    avgNormRes += normRes;
                    resCounter++;
                }
            }
            if (resCounter == 0) {
                resCounter = 1;
                avgRes = 0;
                avgNormRes = 0;
            }
            if (((avgNormRes / resCounter < prm.rdf_rel_tol || avgRes / resCounter < prm.rdf_tol) && iter >= 1) ||
                iter + 1 == prm.rdf_iterations)
                break;
        }
    }

    // -------------------------------------------------------------- A3-A5 ---
    // reconstruction::mapAlphaField (reconstruction.C:725-784), the part below the mesh.changing() && mapAlphaField_ test
    double alphaMappingTime = 0;
    void mapAlphaField(scalar lowerRefineLevel, scalar upperRefineLevel)
    {
        for (label celli = 0; celli < mesh.nCells; ++celli) {
            if (alpha[celli] >= lowerRefineLevel && alpha[celli] <= upperRefineLevel) {
                cutCell_->calcSubCell(celli, interfaceN[celli], interfaceD[celli], false);
                alpha[celli] = cutCell_->volumeOfFluid();
            }
        }
        correctAlphaBCs();
        alphaOld = alpha;
    }

    void reconstruct()  // reconstruction.C:680-722
    {
        initialize();
        Un0.assign(mixedCells.size(), 0.0);
        if (mixedCells.empty()) return;
        if (prm.orientation_method == SVOF_ORIENT_ALPHA_GRAD) calcInterfaceNFromRegAlphaGrad();  // reconstruction.C:694-713
        else if (prm.orientation_method == SVOF_ORIENT_ISO_RDF) calcInterfaceNFromIsoRDF();
        else calcInterfaceNFromIsoAlphaGrad();
        for (size_t i = 0; i < mixedCells.size(); ++i) {
            const label c = mixedCells[i];
            cellStatus[i] = cutCell_->findSignedDistance(c, alpha[c], interfaceN[c], prm.split_warped_face != 0,
                                                         interfaceD[c], interfaceC[c], interfaceS[c]);
        }
    }

    // ---------------------------------------------------------------- A15 ---
    // volPointInterpolation::interpolate(U) evaluated lazily at one point
    // (makeInternalWeights / makeBoundaryWeights / interpolate*Field; OF, recalled)
    vec pointU(label p) const
    {
        const point& pt = mesh.points[p];
        if (!mesh.isPatchPoint[p]) {
            const label n = mesh.pointCells.size(p);
            const label* pc = mesh.pointCells.row(p);
            scalar sumW = 0.0;
            for (label k = 0; k < n; ++k) sumW += 1.0 / mag(pt - mesh.C[pc[k]]);
            vec val;
            for (label k = 0; k < n; ++k) {
                const scalar pw = (1.0 / mag(pt - mesh.C[pc[k]])) / sumW;
                val += pw * U[pc[k]];
            }
            return val;
        }
        const label n = mesh.pointBFaces.size(p);
        const label* pf = mesh.pointBFaces.row(p);
        scalar sumW = 0.0;
        for (label k = 0; k < n; ++k)
            if (mesh.isPatchFace[pf[k]]) sumW += 1.0 / mag(pt - mesh.Cf[mesh.nInternalFaces + pf[k]]);
        vec val;
        for (label k = 0; k < n; ++k) {
            if (!mesh.isPatchFace[pf[k]]) continue;
            const scalar pw = (1.0 / mag(pt - mesh.Cf[mesh.nInternalFaces + pf[k]])) / sumW;
            val += pw * Ub[pf[k]];
        }
        return val;
    }

    // tetrahedron::pointToBarycentric (OF, recalled).  a=cell centre, b,c,d = tri.
    static scalar pointToBarycentric(const point& a, const point& b, const point& c, const point& d, const point& pt,
                                     scalar bary[4])
    {
        const vec v0(a - d), v1(b - d), v2(c - d);
        // tensor t(v0.x v1.x v2.x; v0.y v1.y v2.y; v0.z v1.z v2.z)
        const scalar xx = v0.x, xy = v1.x, xz = v2.x, yx = v0.y, yy = v1.y, yz = v2.y, zx = v0.z, zy = v1.z, zz = v2.z;
        const scalar detT = (xx * yy * zz + xy * yz * zx + xz * yx * zy - xx * yz * zy - xy * yx * zz - xz * yy * zx);
        if (mag(detT) < SMALL) {
            bary[0] = bary[1] = bary[2] = bary[3] = 0.25;
            return detT;
        }
        // inv(t, detT) = cofactor-transpose / detT
        const scalar ixx = (yy * zz - zy * yz) / detT, ixy = (xz * zy - xy * zz) / detT, ixz = (xy * yz - xz * yy) / detT;
        const scalar iyx = (zx * yz - yx * zz) / detT, iyy = (xx * zz - xz * zx) / detT, iyz = (yx * xz - xx * yz) / detT;
        const scalar izx = (yx * zy - yy * zx) / detT, izy = (xy * zx - xx * zy) / detT, izz = (xx * yy - yx * xy) / detT;
        const vec r(pt - d);
        const scalar rx = ixx * r.x + ixy * r.y + ixz * r.z;
        const scalar ry = iyx * r.x + iyy * r.y + iyz * r.z;
        const scalar rz = izx * r.x + izy * r.y + izz * r.z;
        bary[0] = rx;
        bary[1] = ry;
        bary[2] = rz;
        bary[3] = 1 - (rx + ry + rz);
        return detT;
    }

    // interpolationCellPoint<vector>::interpolate(position, celli)
    // = cellPointWeight::findTetrahedron + weighted sum (OF, recalled)
    vec interpolateU(const point& position, label celli) const
    {
        const scalar tol = SMALL;  // cellPointWeight::tol
        const scalar cellVolume = mesh.V[celli];
        const point& cc = mesh.C[celli];
        scalar w[4];
        label tri[3];
        auto tetTri = [&](label f, label tetPt, label out[3]) {  // tetIndices::faceTriIs
            const label* fp = mesh.faces.row(f);
            const label n = mesh.faces.size(f);
            const label base = mesh.tetBasePt[f];
            label facePtI = (tetPt + base) % n;
            label faceOtherPtI = (facePtI + 1) % n;
            if (mesh.owner[f] != celli) std::swap(facePtI, faceOtherPtI);
            out[0] = fp[base];
            out[1] = fp[facePtI];
            out[2] = fp[faceOtherPtI];
        };
        bool found = false;
        const label* cf = mesh.cells.row(celli);
        for (label fi = 0; fi < mesh.cells.size(celli) && !found; ++fi) {
            const label f = cf[fi];
            for (label tetPt = 1; tetPt < mesh.faces.size(f) - 1 && !found; ++tetPt) {
                tetTri(f, tetPt, tri);
                const scalar det = pointToBarycentric(cc, mesh.points[tri[0]], mesh.points[tri[1]], mesh.points[tri[2]],
                                                      position, w);
                if (mag(det / cellVolume) > tol) {
                    const scalar u = w[0], v = w[1], ww = w[2];
                    if ((u + tol > 0) && (v + tol > 0) && (ww + tol > 0) && (u + v + ww < 1 + tol)) found = true;
                }
            }
        }
        if (!found) {
            // nearest tet: by construction interfaceC lies inside its cell, so this
            // branch is a safety net; use the tet whose clamped barycentrics are closest.
            scalar best = VGREAT;
            label bestTri[3] = {0, 0, 0};
            scalar bw[4] = {0.25, 0.25, 0.25, 0.25};
            for (label fi = 0; fi < mesh.cells.size(celli); ++fi) {
                const label f = cf[fi];
                for (label tetPt = 1; tetPt < mesh.faces.size(f) - 1; ++tetPt) {
                    label t3[3];
                    scalar tw[4];
                    tetTri(f, tetPt, t3);
                    pointToBarycentric(cc, mesh.points[t3[0]], mesh.points[t3[1]], mesh.points[t3[2]], position, tw);
                    scalar viol = 0;
                    for (int q = 0; q < 4; ++q) viol += (tw[q] < 0) ? -tw[q] : 0;
                    if (viol < best) {
                        best = viol;
                        std::memcpy(bestTri, t3, sizeof(t3));
                        std::memcpy(bw, tw, sizeof(tw));
                    }
                }
            }
            std::memcpy(tri, bestTri, sizeof(tri));
            std::memcpy(w, bw, sizeof(w));
        }
        vec t = U[celli] * w[0];
        t += pointU(tri[0]) * w[1];
        t += pointU(tri[1]) * w[2];
        t += pointU(tri[2]) * w[3];
        return t;
    }

    // advectionTemplates.C:40-75 (empty patches return zero)
    bool faceActive(label f) const
    {
        if (f < mesh.nInternalFaces) return true;
        const svof_patch& pt = mesh.patches[mesh.patchID[f - mesh.nInternalFaces]];
        return pt.kind != SVOF_PATCH_EMPTY && pt.size > 0;
    }
    scalar faceValue(const std::vector<scalar>& fld, label f) const { return faceActive(f) ? fld[f] : 0.0; }
    void setFaceValue(std::vector<scalar>& fld, label f, scalar v) const
    {
        if (faceActive(f)) fld[f] = v;
    }

    // ---------------------------------------------------------------- A7 ----
    void timeIntegratedFlux(scalar dt)  // advection.C:85-221
    {
        std::vector<label> bsFaces;
        std::vector<vec> bsn0;
        std::vector<scalar> bsD0, bsUn0;
        for (size_t i = 0; i < mixedCells.size(); ++i) {
            if (cellStatus[i] != 0) continue;
            const label cellI = mixedCells[i];
            const vec& normalI = interfaceN[cellI];
            const scalar distanceI = interfaceD[cellI];
            const point& centreI = interfaceC[cellI];
            const scalar un0 = interpolateU(centreI, cellI) & normalI;
            Un0[i] = un0;
            const label* cf = mesh.cells.row(cellI);
            for (label k = 0; k < mesh.cells.size(cellI); ++k) {
                const label facei = cf[k];
                if (mesh.isInternalFace(facei)) {
                    bool isDownwindFace = false;
                    if (cellI == mesh.owner[facei]) {
                        if (phi[facei] >= 0.0) isDownwindFace = true;
                    } else {
                        if (phi[facei] < 0.0) isDownwindFace = true;
                    }
                    if (isDownwindFace) {
                        dVf[facei] = cutFace_->timeIntegratedFaceFlux(facei, normalI, distanceI, un0, dt, phi[facei],
                                                                      mesh.magSf[facei]);
                    }
                } else {
                    bsFaces.push_back(facei);
                    bsn0.push_back(normalI);
                    bsD0.push_back(distanceI);
                    bsUn0.push_back(un0);
                }
            }
        }
        for (size_t i = 0; i < bsFaces.size(); ++i) {
            const label faceI = bsFaces[i];
            if (!faceActive(faceI)) continue;  // phib[patchI].empty()
            const scalar phiP = phi[faceI];
            if (phiP >= 0.0) {
                dVf[faceI] =
                    cutFace_->timeIntegratedFaceFlux(faceI, bsn0[i], bsD0[i], bsUn0[i], dt, phiP, mesh.magSf[faceI]);
            }
        }
    }

    // advection.C:224-256
    void setDownwindFaces(label cellI, std::vector<label>& downwindFaces) const
    {
        downwindFaces.clear();
        const label* c = mesh.cells.row(cellI);
        for (label k = 0; k < mesh.cells.size(cellI); ++k) {
            const label facei = c[k];
            const scalar ph = faceValue(phi, facei);
            if (mesh.owner[facei] == cellI) {
                if (ph >= 0) downwindFaces.push_back(facei);
            } else if (ph < 0) {
                downwindFaces.push_back(facei);
            }
        }
    }

    // advection.C:259-288
    scalar netFlux(label cellI, const std::vector<scalar>& fld) const
    {
        scalar dV = 0.0;
        const label* c = mesh.cells.row(cellI);
        for (label k = 0; k < mesh.cells.size(cellI); ++k) {
            const label facei = c[k];
            const scalar dVff = faceValue(fld, facei);
            if (mesh.owner[facei] == cellI) {
                dV += dVff;
            } else {
                dV -= dVff;
            }
        }
        return dV;
    }

    // advection.C:54-82
    void extendMarkedCells(std::vector<char>& markedCell) const
    {
        std::vector<char> markedFace(mesh.nFaces, 0);
        for (label c = 0; c < mesh.nCells; ++c)
            if (markedCell[c])
                for (label k = 0; k < mesh.cells.size(c); ++k) markedFace[mesh.cells.row(c)[k]] = 1;
        for (label f = 0; f < mesh.nInternalFaces; ++f)
            if (markedFace[f]) {
                markedCell[mesh.owner[f]] = 1;
                markedCell[mesh.neighbour[f]] = 1;
            }
        for (label f = mesh.nInternalFaces; f < mesh.nFaces; ++f)
            if (markedFace[f]) markedCell[mesh.owner[f]] = 1;
    }

    // advectionTemplates.C:218-349
    void boundFlux(const std::vector<char>& nextToInterface, std::vector<scalar>& dVfCorrectionValues,
                   std::vector<label>& correctedFaces, const scalar* Sp, const scalar* Su, scalar dt)
    {
        const scalar rDeltaT = 1.0 / dt;
        correctedFaces.clear();
        std::vector<label> downwindFaces, facesToPassFluidThrough;
        std::vector<scalar> dVfmax, phiL;
        for (label celli = 0; celli < mesh.nCells; ++celli) {
            if (!nextToInterface[celli]) continue;
            if (alpha[celli] < -aTol || alpha[celli] > 1.0 + aTol) {
                const scalar Vi = mesh.V[celli];
                scalar alphaOvershoot = pos0(alpha[celli] - 1.0) * (alpha[celli] - 1.0) + neg0(alpha[celli]) * alpha[celli];
                scalar fluidToPassOn = alphaOvershoot * Vi;
                label nFacesToPassFluidThrough = 1;
                bool firstLoop = true;
                for (label iter = 0; iter < 10; iter++) {
                    if (mag(alphaOvershoot) < aTol || nFacesToPassFluidThrough == 0) break;
                    facesToPassFluidThrough.clear();
                    dVfmax.clear();
                    phiL.clear();
                    setDownwindFaces(celli, downwindFaces);
                    scalar dVftot = 0;
                    nFacesToPassFluidThrough = 0;
                    for (const label facei : downwindFaces) {
                        const scalar phif = faceValue(phi, facei);
                        const scalar dVff = faceValue(dVf, facei) + faceValue(dVfCorrectionValues, facei);
                        const scalar maxExtraFaceFluidTrans = mag(pos0(fluidToPassOn) * phif * dt - dVff);
                        if (maxExtraFaceFluidTrans / Vi > aTol) {
                            facesToPassFluidThrough.push_back(facei);
                            phiL.push_back(phif);
                            dVfmax.push_back(maxExtraFaceFluidTrans);
                            dVftot += mag(phif * dt);
                        }
                    }
                    for (size_t fi = 0; fi < facesToPassFluidThrough.size(); ++fi) {
                        const label facei = facesToPassFluidThrough[fi];
                        scalar fluidToPassThroughFace = mag(fluidToPassOn) * mag(phiL[fi] * dt) / dVftot;
                        nFacesToPassFluidThrough += label(pos0(dVfmax[fi] - fluidToPassThroughFace));
                        fluidToPassThroughFace = smin(fluidToPassThroughFace, dVfmax[fi]);
                        scalar dVff = faceValue(dVfCorrectionValues, facei);
                        dVff += sign(phiL[fi]) * sign(fluidToPassOn) * fluidToPassThroughFace;
                        setFaceValue(dVfCorrectionValues, facei, dVff);
                        if (firstLoop) correctedFaces.push_back(facei);
                    }
                    firstLoop = false;
                    const scalar SuI = Su ? Su[celli] : 0.0, SpI = Sp ? Sp[celli] : 0.0;
                    scalar alpha1New = (alphaOld[celli] * rDeltaT + SuI - netFlux(celli, dVf) / Vi * rDeltaT -
                                        netFlux(celli, dVfCorrectionValues) / Vi * rDeltaT) /
                                       (rDeltaT - SpI);
                    alphaOvershoot = pos0(alpha1New - 1.0) * (alpha1New - 1.0) + neg0(alpha1New) * alpha1New;
                    fluidToPassOn = alphaOvershoot * Vi;
                }
            }
        }
    }

    static void minMax(const std::vector<scalar>& a, scalar& mn, scalar& mx)
    {
        mn = VGREAT;
        mx = -VGREAT;
        for (scalar v : a) {
            mn = smin(mn, v);
            mx = smax(mx, v);
        }
    }

    // advectionTemplates.C:118-215
    void limitFlux(const scalar* Sp, const scalar* Su, scalar dt)
    {
        scalar mx, mn;
        minMax(alpha, mn, mx);
        scalar maxAlphaMinus1 = mx - 1.0;
        scalar minAlpha = mn;
        minBefore = minAlpha;
        maxM1Before = maxAlphaMinus1;
        nSweeps = 0;

        std::vector<scalar> dVfCorrectionValues(mesh.nFaces, 0.0);
        std::vector<char> needBounding(mesh.nCells, 0);
        for (label c : mixedCells) needBounding[c] = 1;
        extendMarkedCells(needBounding);

        for (label n = 0; n < prm.n_alpha_bounds; n++) {
            if (maxAlphaMinus1 > aTol || minAlpha < -aTol) {
                std::vector<label> correctedFaces;
                std::fill(dVfCorrectionValues.begin(), dVfCorrectionValues.end(), 0.0);
                boundFlux(needBounding, dVfCorrectionValues, correctedFaces, Sp, Su, dt);
                std::vector<char> alreadyUpdated(mesh.nFaces, 0);
                for (const label facei : correctedFaces) {
                    if (!alreadyUpdated[facei]) {
                        alreadyUpdated[facei] = 1;
                        const label own = mesh.owner[facei];
                        scalar Vown = mesh.V[own];
                        alpha[own] -= faceValue(dVfCorrectionValues, facei) / Vown;
                        if (mesh.isInternalFace(facei)) {
                            const label nei = mesh.neighbour[facei];
                            scalar Vnei = mesh.V[nei];
                            alpha[nei] += faceValue(dVfCorrectionValues, facei) / Vnei;
                        }
                        const scalar corrVf = faceValue(dVf, facei) + faceValue(dVfCorrectionValues, facei);
                        setFaceValue(dVf, facei, corrVf);
                    }
                }
                nSweeps++;
            } else {
                break;
            }
            minMax(alpha, mn, mx);
            maxAlphaMinus1 = mx - 1.0;
            minAlpha = mn;
        }
        minAfter = minAlpha;
        maxM1After = maxAlphaMinus1;
        correctAlphaBCs();
    }

    // advection.C:291-308
    void applyBruteForceBounding()
    {
        if (prm.snap_tol > 0.0) {
            const scalar tol = prm.snap_tol;
            for (scalar& a : alpha) a = a * pos0(a - tol) * neg0(a - (1.0 - tol)) + pos0(a - (1.0 - tol));
            correctAlphaBCs();
        }
        if (prm.clip) {
            for (scalar& a : alpha) a = smin(scalar(1), smax(scalar(0), a));
            correctAlphaBCs();
        }
    }

    // ------------------------------------------------------------ A6,A10 ----
    void advect(scalar dt, const scalar* Sp, const scalar* Su)  // advectionTemplates.C:352-418
    {
        const scalar rDeltaT = 1.0 / dt;
        const label nIF = mesh.nInternalFaces;
        alphaOld = alpha;  // alpha1.oldTime() (SURVEY 8a' item 19)

        // dVf_ = upwind<scalar>(mesh_, phi_).flux(alpha1_) * deltaT   (:371)
        for (label f = 0; f < nIF; ++f) {
            const scalar af = (phi[f] >= 0) ? alpha[mesh.owner[f]] : alpha[mesh.neighbour[f]];
            dVf[f] = (phi[f] * af) * dt;
        }
        for (label f = nIF; f < mesh.nFaces; ++f) dVf[f] = faceActive(f) ? (phi[f] * alphaB[f - nIF]) * dt : 0.0;

        timeIntegratedFlux(dt);  // :374

        // alpha = (alphaOld*rDeltaT + Su - surfaceIntegrate(dVf)*rDeltaT)/(rDeltaT - Sp)   (:399-404)
        std::vector<scalar> ivf(mesh.nCells, 0.0);
        for (label f = 0; f < nIF; ++f) {
            ivf[mesh.owner[f]] += dVf[f];
            ivf[mesh.neighbour[f]] -= dVf[f];
        }
        for (label f = nIF; f < mesh.nFaces; ++f)
            if (faceActive(f)) ivf[mesh.owner[f]] += dVf[f];
        for (label c = 0; c < mesh.nCells; ++c) {
            ivf[c] /= mesh.V[c];
            scalar num = alphaOld[c] * rDeltaT;
            if (Su) num = num + Su[c];
            num = num - ivf[c] * rDeltaT;
            alpha[c] = num / (Sp ? (rDeltaT - Sp[c]) : rDeltaT);
        }
        correctAlphaBCs();  // :406

        limitFlux(Sp, Su, dt);      // :409
        applyBruteForceBounding();  // :413

        for (label f = 0; f < mesh.nFaces; ++f) alphaPhi[f] = dVf[f] / dt;  // :417
    }

    scalar volume() const  // plicVof.H:44-46  gSum(alpha*V)
    {
        scalar s = 0;
        for (label c = 0; c < mesh.nCells; ++c) s += alpha[c] * mesh.V[c];
        return s;
    }
};

}  // namespace ora
