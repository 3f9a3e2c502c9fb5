// svof_plic_warp.cuh -- plane positioning (cutCell::findSignedDistance, cutCell.C:611-799) with SELF-CONTAINED
// groups of 8 lanes per mixed cell: four cells per warp, no CTA barrier anywhere.
//
// Round 1's k_plic_group alternated "face phases" (all lanes) and "leader phases" (one lane per cell, packed into one
// warp) with six __syncthreads per calcSubCell evaluation; ncu showed 23.7 % of the warp slots active, 12.7 of 32
// lanes per instruction and the CTA barrier as the top stall.  Here every lane of a group executes the ORDERED
// reductions the reference performs sequentially (interface polygon centre, segment sums, pyramid volumes, the
// bracket / cubic / Newton search state) redundantly from shared memory -- the FP64 pipe was at 22 %, so the
// redundancy is free -- and the lanes only meet at three __syncwarp(groupMask) per evaluation.  Results are bitwise
// those of the thread-per-cell form (and of the oracle): same operands in the same order.
//
// The per-cell staging block in shared memory is sized at RUN TIME from the mesh's maxima (local faces per cell,
// vertices per face, local points per cell), not from the capacity variant's caps: a 14-face Kelvin cell takes 8.5 KB
// instead of the 18 KB the 40-face cap reserved.
//   res[k]   19 doubles: sub-face centre(3), area(3), up to 4 interface points(12), {status, nip}
//   seg[k]   21 doubles: up to 3 interface segments x { a*(p0+p1+fC) (3), (p1-p0)^(fC-p0) (3), a }
//   pv[k]    pyramid 3*volume of the sub-face
//   rec[k]   4*maxFV+6 doubles: vertices, n.p per vertex, whole-face centre/area (plane independent)
//   vd, vdRaw sorted / unsorted vertex distances;  lfFace, fnv, lfTri: the local face list
#pragma once
#include "svof_plic_group.cuh"

namespace svof {

#ifndef SV_PW_THREADS
#define SV_PW_THREADS 128
#define SV_PW_MINB 4
#endif

struct PlicWarpLayout {
    int mLF, maxFV, mLP, recD, strideD;  // strideD: doubles per cell, == 4 (mod 16) so the four cells of a warp hit distinct banks
    __host__ __device__ static PlicWarpLayout make(int mLF, int maxFV, int mLP)
    {
        PlicWarpLayout L;
        L.mLF = mLF;
        L.maxFV = maxFV;
        L.mLP = mLP;
        L.recD = 4 * maxFV + 6;
        int d = mLF * (19 + 21 + 1 + L.recD) + 2 * mLP + (mLF * 10 + 7) / 8;  // ints: lfFace, fnv (4 B), lfTri (2 B)
        while ((d & 15) != 4) ++d;
        L.strideD = d;
        return L;
    }
};

template <class CP>
__device__ __forceinline__ void plicCellWarp(const MeshDev& m, const PlicWarpLayout& L, double* blk, int g, unsigned gmask, int grpLane0,
                                             int i, const int* mixedCells, const double* __restrict__ alpha, const double* iN, bool splitB,
                                             int* cellStatus, double* iD, double* iC, double* iS, int& err)
{
    double* res = blk;
    double* seg = res + L.mLF * 19;
    double* pv = seg + L.mLF * 21;
    double* frec = pv + L.mLF;
    double* vd = frec + L.mLF * L.recD;
    double* vdRaw = vd + L.mLP;
    int* lfFace = reinterpret_cast<int*>(vdRaw + L.mLP);
    int* fnv = lfFace + L.mLF;
    short* lfTri = reinterpret_cast<short*>(fnv + L.mLF);

    const int cell = mixedCells[i];
    const double alphaI = alpha[cell];
    const d3 nL = ld3(iN, cell);
    const double Vcell = __ldg(m.V + cell);
    const int c0 = __ldg(m.cellOff + cell), c1 = __ldg(m.cellOff + cell + 1);
    const int p0 = __ldg(m.cellPtOff + cell), p1 = __ldg(m.cellPtOff + cell + 1);
    int nl = 0, nP = 0;
    // ---- stage the polyhedron: local face list and vertex distances (cutCell.C:637-662; :140-236 when split)
    if (!splitB) {
        nl = c1 - c0;
        if (nl > L.mLF) { err |= SVERR_CELL_FACES; nl = L.mLF; }
        for (int k = g; k < nl; k += SV_G) {
            lfFace[k] = __ldg(m.cellFaces + c0 + k);
            lfTri[k] = -1;
        }
        nP = p1 - p0;
        if (nP > L.mLP) { err |= SVERR_CELL_POINTS; nP = L.mLP; }
        for (int k = g; k < nP; k += SV_G) vdRaw[k] = -dot(nL, ld3(m.points, __ldg(m.cellPts + p0 + k)));
    } else {
        if (g == 0) {  // the triangulated local polyhedron is enumerated in order by one lane (rare configuration)
            for (int k = c0; k < c1; ++k) {
                const int f = __ldg(m.cellFaces + k);
                if (!(__ldg(m.flat + f) > (1.0 - SV_TSMALL))) {
                    const int nv = __ldg(m.faceOff + f + 1) - __ldg(m.faceOff + f);
                    for (int t = 0; t < nv; ++t) {
                        if (nl < L.mLF) { lfFace[nl] = f; lfTri[nl] = (short)t; nl++; } else err |= SVERR_CELL_FACES;
                    }
                } else {
                    if (nl < L.mLF) { lfFace[nl] = f; lfTri[nl] = -1; nl++; } else err |= SVERR_CELL_FACES;
                }
            }
            for (int k = p0; k < p1; ++k) {
                if (nP < L.mLP) vdRaw[nP++] = -dot(nL, ld3(m.points, __ldg(m.cellPts + k))); else err |= SVERR_CELL_POINTS;
            }
            for (int k = c0; k < c1; ++k) {
                const int f = __ldg(m.cellFaces + k);
                if (!(__ldg(m.flat + f) > (1.0 - SV_TSMALL))) {
                    if (nP < L.mLP) vdRaw[nP++] = -dot(nL, ld3(m.Cf, f)); else err |= SVERR_CELL_POINTS;
                }
            }
        }
        nl = __shfl_sync(gmask, nl, grpLane0);
        nP = __shfl_sync(gmask, nP, grpLane0);
    }
    __syncwarp(gmask);
    // stable descending sort by rank (values only: ties are indistinguishable, cutCell.C:664-670)
    for (int k = g; k < nP; k += SV_G) {
        const double v = vdRaw[k];
        int rank = 0;
        for (int j = 0; j < nP; ++j) {
            const double w = vdRaw[j];
            rank += ((w > v) || (w == v && j < k)) ? 1 : 0;
        }
        vd[rank] = v;
    }
    // plane-independent face records
    for (int k = g; k < nl; k += SV_G) {
        d3 fp[CP::MAXFV];
        const int nv = loadLocalFace<CP>(m, cell, lfFace[k], lfTri[k], splitB, fp, err);
        FaceRec rec{frec + (size_t)k * L.recD, L.maxFV};
        for (int q = 0; q < nv; ++q) {
            rec.setFp(q, fp[q]);
            rec.pn(q) = dot(fp[q], nL);
        }
        rec.setFull(faceCentreOF(fp, nv), faceAreaNormalOF(fp, nv));
        fnv[k] = nv;
    }
    __syncwarp(gmask);

    // ---- search state (cutCell.C:682-719), identical on every lane of the group
    double lowDistance = vd[0], upDistance = vd[nP - 1], lowAlpha = 0.0, upAlpha = 1.0;
    int lowLabel = 0, upLabel = nP - 1;
    double midLabel = 0, aOneThird = 0, deltaDistance = 0, curD = 0;
    int phase = 0;  // 0 bracketing, 1 collapsed bracket, 2 one third, 3 two thirds, 4 final
    int outStatus = 0;
    bool active = true;
    if (mag(nL) < SV_TSMALL) {  // cutCell.C:625-628: D/C/S stay untouched
        outStatus = int(sgn(0.5 - alphaI));
        active = false;
    } else if ((upLabel - lowLabel) > 1) {
        midLabel = round(0.5 * (upLabel + lowLabel));
        curD = vd[int(midLabel)];
        phase = 0;
    } else if (fabs(lowDistance - upDistance) < SV_TSMALL) {
        curD = 0.5 * (lowDistance + upDistance);
        phase = 1;
    } else {
        deltaDistance = (upDistance - lowDistance) / 3.0;
        curD = lowDistance + deltaDistance;
        phase = 2;
    }

    // ---- one calcSubCell (cutCell.C:343-542) per iteration
    while (active) {
        // A. every lane clips its faces
        for (int k = g; k < nl; k += SV_G) {
            double* r = res + 19 * k;
            const FaceRec rec{frec + (size_t)k * L.recD, L.maxFV};
            d3 c, a;
            int nip;
            const int st = clipFaceStream<CP>(rec, fnv[k], curD, rec.fullC(), rec.fullA(), c, a, reinterpret_cast<d3*>(r + 6), nip, err);
            r[0] = c.x; r[1] = c.y; r[2] = c.z;
            r[3] = a.x; r[4] = a.y; r[5] = a.z;
            *reinterpret_cast<int2*>(r + 18) = make_int2(st, nip);
        }
        __syncwarp(gmask);
        // B. classification + mean of the interface edge points (cutCell.C:37-54), all lanes
        bool fullySubmerged = true, fullyEmpty = true;
        int nSubmergedFaces = 0, nCut = 0;
        d3 fC = zero3();
        {
            int nEp = 0;
            for (int k = 0; k < nl; ++k) {
                const double* r = res + 19 * k;
                const int2 sn = *reinterpret_cast<const int2*>(r + 18);
                if (sn.x == 0) {
                    fullySubmerged = false;
                    fullyEmpty = false;
                    nCut++;
                    for (int q = 0; q < sn.y; ++q) {
                        fC += mk3(r[6 + 3 * q], r[7 + 3 * q], r[8 + 3 * q]);
                        nEp++;
                    }
                } else if (sn.x == -1) {
                    fullyEmpty = false;
                    nSubmergedFaces++;
                    nCut++;
                } else {
                    fullySubmerged = false;
                }
            }
            if (nEp > 0) fC /= double(nEp);
        }
        const bool cutAny = !fullySubmerged && !fullyEmpty;
        // C. lanes: interface segments of their cut faces (cutCell.C:60-79, the per-segment part)
        if (cutAny) {
            for (int k = g; k < nl; k += SV_G) {
                const double* r = res + 19 * k;
                const int2 sn = *reinterpret_cast<const int2*>(r + 18);
                if (sn.x != 0) continue;
                double* s = seg + 21 * k;
                for (int pi = 0; pi < sn.y - 1; ++pi) {
                    const d3 q0 = mk3(r[6 + 3 * pi], r[7 + 3 * pi], r[8 + 3 * pi]);
                    const d3 nx = mk3(r[9 + 3 * pi], r[10 + 3 * pi], r[11 + 3 * pi]);
                    const d3 c = q0 + nx + fC;
                    const d3 nn = cross(nx - q0, fC - q0);
                    const double a = mag(nn);
                    const d3 ac = a * c;
                    s[7 * pi] = ac.x; s[7 * pi + 1] = ac.y; s[7 * pi + 2] = ac.z;
                    s[7 * pi + 3] = nn.x; s[7 * pi + 4] = nn.y; s[7 * pi + 5] = nn.z;
                    s[7 * pi + 6] = a;
                }
            }
        }
        __syncwarp(gmask);
        // D. ordered accumulation, interface centre/area, cEst (cutCell.C:56-99,110), all lanes
        int status = 0;
        double VOF = 0.0;
        bool needVolume = false;
        d3 iCl = zero3(), iSl = zero3(), cEst = zero3();
        if (cutAny) {
            d3 sumN = zero3(), sumAc = zero3();
            double sumA = 0.0;
            for (int k = 0; k < nl; ++k) {
                const int2 sn = *reinterpret_cast<const int2*>(res + 19 * k + 18);
                if (sn.x != 0) continue;
                const double* s = seg + 21 * k;
                for (int pi = 0; pi < sn.y - 1; ++pi) {
                    const d3 nn = mk3(s[7 * pi + 3], s[7 * pi + 4], s[7 * pi + 5]);
                    sumN += sgn(dot(nn, sumN)) * nn;
                    sumA += s[7 * pi + 6];
                    sumAc += mk3(s[7 * pi], s[7 * pi + 1], s[7 * pi + 2]);
                }
            }
            if (sumA < SV_ROOTVSMALL) {
                iCl = fC;
                iSl = zero3();
            } else {
                iCl = (1.0 / 3.0) * sumAc / sumA;
                iSl = 0.5 * sumN;
            }
            if (dot(iSl, iCl - zero3()) < 0.0) iSl = iSl * (-1.0);  // vs the origin: SURVEY 8a' item 24
            if (mag(iSl) < SV_TSMALL) {
                if (nSubmergedFaces == 0) { status = 1; VOF = 0.0; } else { status = -1; VOF = 1.0; }
            } else {
                status = 0;
                needVolume = true;
                for (int k = 0; k < nl; ++k) {
                    const double* r = res + 19 * k;
                    if (*reinterpret_cast<const int*>(r + 18) <= 0) cEst += mk3(r[0], r[1], r[2]);
                }
                cEst += iCl;
                cEst /= double(nCut + 1);
            }
        } else if (fullyEmpty) {
            status = 1;
            VOF = 0.0;
        } else {
            status = -1;
            VOF = 1.0;
        }
        // E. lanes: pyramid volumes of their sub-faces (cutCell.C:116-123)
        if (needVolume) {
            for (int k = g; k < nl; k += SV_G) {
                const double* r = res + 19 * k;
                if (*reinterpret_cast<const int*>(r + 18) <= 0)
                    pv[k] = dmax(fabs(dot(mk3(r[3], r[4], r[5]), mk3(r[0], r[1], r[2]) - cEst)), SV_VSMALL);
            }
        }
        __syncwarp(gmask);
        // F. ordered volume sum, then advance the search (cutCell.C:691-799), all lanes
        if (needVolume) {
            double vol = 0.0;
            for (int k = 0; k < nl; ++k)
                if (*reinterpret_cast<const int*>(res + 19 * k + 18) <= 0) vol += pv[k];
            vol += dmax(fabs(dot(iSl, iCl - cEst)), SV_VSMALL);
            vol /= 3.0;
            VOF = vol / Vcell;
        }
        bool finish = false;
        if (phase == 0) {
            const double midAlpha = VOF;
            if (fabs(midAlpha - alphaI) < SV_TSMALL) {
                finish = true;
            } else {
                if (midAlpha > alphaI) {
                    upLabel = int(midLabel);
                    upDistance = curD;
                    upAlpha = midAlpha;
                } else {
                    lowLabel = int(midLabel);
                    lowDistance = curD;
                    lowAlpha = midAlpha;
                }
                if ((upLabel - lowLabel) > 1) {
                    midLabel = round(0.5 * (upLabel + lowLabel));
                    curD = vd[int(midLabel)];
                } else if (fabs(lowDistance - upDistance) < SV_TSMALL) {
                    curD = 0.5 * (lowDistance + upDistance);
                    phase = 1;
                } else {
                    deltaDistance = (upDistance - lowDistance) / 3.0;
                    curD = lowDistance + deltaDistance;
                    phase = 2;
                }
            }
        } else if (phase == 1 || phase == 4) {
            finish = true;
        } else if (phase == 2) {
            aOneThird = VOF - lowAlpha;
            curD = lowDistance + 2.0 * deltaDistance;
            phase = 3;
        } else {  // phase 3: cubic + Newton (cutCell.C:744-789)
            const double alphaPrismatoid = upAlpha - lowAlpha;
            const double alphaOneThird = aOneThird;
            const double alphaTwoThirds = VOF - lowAlpha;
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
            curD = lowDistance - lambda * (lowDistance - upDistance);
            phase = 4;
        }
        if (finish) {
            outStatus = status;
            if (g == 0) {
                iD[cell] = curD;
                st3(iC, cell, iCl);
                st3(iS, cell, iSl);
            }
            active = false;
        }
    }
    if (g == 0) cellStatus[i] = outStatus;
}

// Persistent warps pull batches of four cells from an atomic counter (ctl->plicNext, reset by k_ctl_reset_recon).
template <class CP>
__global__ void __launch_bounds__(SV_PW_THREADS, SV_PW_MINB) k_plic_warp(MeshDev m, PlicWarpLayout L, const int* mixedCells, Ctl* ctl,
                                                                        const double* __restrict__ alpha, const double* iN, int split,
                                                                        int* cellStatus, double* iD, double* iC, double* iS)
{
    extern __shared__ __align__(16) double smemPW[];
    const int lane = threadIdx.x & 31, g = lane & (SV_G - 1), grp = lane >> 3;
    const unsigned gmask = 0xFFu << (grp * SV_G);
    double* blk = smemPW + (size_t)((threadIdx.x >> 5) * 4 + grp) * L.strideD;
    const int nMixed = ctl->nMixed;
    int err = 0;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&ctl->plicNext, 1) * 4;
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= nMixed) break;
        const int i = base + grp;
        if (i < nMixed) plicCellWarp<CP>(m, L, blk, g, gmask, grp * SV_G, i, mixedCells, alpha, iN, split != 0, cellStatus, iD, iC, iS, err);
        __syncwarp();
    }
    if (err) atomicOr(&ctl->err, err);
}

}  // namespace svof
