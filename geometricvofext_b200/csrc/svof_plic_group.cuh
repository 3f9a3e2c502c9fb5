// svof_plic_group.cuh -- plane positioning (cutCell::findSignedDistance, cutCell.C:611-799) with
// a group of G=8 lanes per mixed cell instead of one thread per cell.
//
// Why: at 256^3 only ~3e4 cells are mixed, i.e. one thread per cell fills <10% of a B200 and each
// thread walks ~35 polygon clips back to back (measured 0.50 ms, FP64 pipe 12% busy, 6 warps/SM).
// Here the faces of a cell are clipped concurrently by the lanes of its group (a hex keeps one face
// per lane entirely in registers, together with the plane-independent parts: n.p per vertex and the
// whole-face centre/area used for submerged faces), and only the ORDERED reductions the reference
// performs sequentially (interface polygon, pyramid volumes: cutCell.C:37-137) are done by the group
// leader, from shared memory, in the reference's order -- so results stay bitwise identical to the
// one-thread version (and to the oracle).
//
// Shared memory per cell (polyhedron staged once per cell):
//   FaceRes res[MAXCF]  sub-face centre/area/status/interface points of each (local) face
//   SegRes  seg[MAXCF]  per-face interface segments (a*c, n, a) once the polygon centre is known
//   vd[MAXCP]           sorted vertex distances;  lf[MAXCF] local face list (splitWarpedFace)
#pragma once
#include "svof_geom.cuh"

namespace svof {

template <class CP>
struct GFaceRes {
    d3 c, a;             // sub-face centre / area vector (status 0 or -1)
    d3 ip[CP::MAXIP];    // interface points (status 0)
    int st, nip;
};
template <class CP>
struct GSegRes {
    d3 ac[CP::MAXIP - 1];  // a * (p0 + p1 + fC)
    d3 nn[CP::MAXIP - 1];  // (p1 - p0) ^ (fC - p0)
    double a[CP::MAXIP - 1];
};
template <class CP>
struct GCellShared {
    GFaceRes<CP> res[CP::MAXCF];
    GSegRes<CP> seg[CP::MAXCF];
    double pv[CP::MAXCF];   // pyramid 3*volumes of the sub-faces
    double vd[CP::MAXCP];
    int lfFace[CP::MAXCF];  // local face -> mesh face
    short lfTri[CP::MAXCF]; // local face -> triangle index of a split warped face, or -1
    // broadcast slots
    double D;
    d3 fC, cEst, iC, iS;
    int nLocal, active, cutAny, nPts;
};

// one local face of the (possibly split) polyhedron into fp[]; returns vertex count
template <class CP>
__device__ __forceinline__ int loadLocalFace(const MeshDev& m, int cell, int f, int tri, bool split, d3* fp, int& err)
{
    int nv = loadFace<CP>(m, f, fp, err);
    if (!split) return nv;
    const bool own = (__ldg(m.owner + f) == cell);
    if (tri < 0) {
        if (!own) reversePoly(fp, nv);
        return nv;
    }
    const int pn = (tri + 1 == nv) ? 0 : tri + 1;
    const d3 a = fp[tri], b = fp[pn];
    fp[0] = ld3(m.Cf, f);
    fp[1] = own ? a : b;
    fp[2] = own ? b : a;
    return 3;
}

// Lane-private scratch in SHARED memory, strided so that element e of thread t sits at base[e * T + t]
// (conflict-free when the lanes of a warp touch the same element).  The first version kept the cached face
// (fp, n.p), the vertex signs and the clipped polygon in thread-local arrays; their dynamic indexing put them
// in local memory, and with 512 threads per SM that working set no longer fits L1: ncu counted 61 M local
// sectors per launch (2 GB through L2), 38% of the local loads and 70% of the stores missing L1, and
// long-scoreboard as the second stall reason after the phase barriers.
template <class CP>
struct LanePriv {
    static constexpr int DOUBLES = 4 * CP::MAXFV;  // fp[MAXFV][3], pn[MAXFV]
    double* base;
    int T;
    __device__ __forceinline__ d3 fp(int q) const { return mk3(base[(3 * q) * T], base[(3 * q + 1) * T], base[(3 * q + 2) * T]); }
    __device__ __forceinline__ void setFp(int q, const d3& v) const
    {
        base[(3 * q) * T] = v.x;
        base[(3 * q + 1) * T] = v.y;
        base[(3 * q + 2) * T] = v.z;
    }
    __device__ __forceinline__ double& pn(int q) const { return base[(3 * CP::MAXFV + q) * T]; }
};

// cutFace::calcSubFace + calcSubFaceCentreAndArea (cutFace.C:37-96,136-259) on the cached face, without
// materialising the clipped polygon: its points are generated twice in the reference's order (first pass: count,
// point sum, first three points; second pass: the edge sums), which gives the same operands in the same order as
// clipFace()/subFaceCentreAndArea() and therefore the same bits.  ip points to shared memory.
// The same face data for cells with more than one face per lane (polyhedra, splitWarpedFace): one record per LOCAL
// FACE of the cell, contiguous, sized from the mesh's actual maxima (not the variant's caps): vertices, n.p per vertex,
// whole-face centre and area.  Without it every evaluation reloaded each face from global memory and clipped it
// through the thread-local path (measured on 14-face Kelvin cells: 0.40 ms for 1.5 k cells).
struct FaceRec {
    double* base;   // [3*maxFV] vertices, [maxFV] n.p, [3] centre, [3] area
    int maxFV;
    static __host__ __device__ int doubles(int maxFV) { return 4 * maxFV + 6; }
    __device__ __forceinline__ d3 fp(int q) const { return mk3(base[3 * q], base[3 * q + 1], base[3 * q + 2]); }
    __device__ __forceinline__ void setFp(int q, const d3& v) const
    {
        base[3 * q] = v.x;
        base[3 * q + 1] = v.y;
        base[3 * q + 2] = v.z;
    }
    __device__ __forceinline__ double& pn(int q) const { return base[3 * maxFV + q]; }
    __device__ __forceinline__ d3 fullC() const { return mk3(base[4 * maxFV], base[4 * maxFV + 1], base[4 * maxFV + 2]); }
    __device__ __forceinline__ d3 fullA() const { return mk3(base[4 * maxFV + 3], base[4 * maxFV + 4], base[4 * maxFV + 5]); }
    __device__ __forceinline__ void setFull(const d3& c, const d3& a) const
    {
        base[4 * maxFV] = c.x; base[4 * maxFV + 1] = c.y; base[4 * maxFV + 2] = c.z;
        base[4 * maxFV + 3] = a.x; base[4 * maxFV + 4] = a.y; base[4 * maxFV + 5] = a.z;
    }
};

template <class CP, class ACC>
__device__ __forceinline__ double liftedS(const ACC& lp, int q, double D)
{
    double si = lp.pn(q) + D;
    if (fabs(si) < SV_TSMALL) si += sgn(si) * SV_TSMALL;
    return si;
}
template <class CP, class ACC>
__device__ __forceinline__ int clipFaceStream(const ACC& lp, int nv, double D, const d3& fullC, const d3& fullA, d3& centre,
                                              d3& area, d3* ip, int& nip, int& err)
{
    int nSub = 0, first = -1;
    for (int i = 0; i < nv; ++i) {
        if (liftedS<CP>(lp, i, D) < 0.0) {
            nSub++;
            if (first < 0) first = i;
        }
    }
    nip = 0;
    if (nSub == nv) {
        centre = fullC;
        area = fullA;
        return -1;
    }
    if (nSub == 0) {
        centre = zero3();
        area = zero3();
        return 1;
    }
    // pass 1
    d3 P0 = zero3(), P1 = zero3(), P2 = zero3(), sum = zero3();
    int np = 0;
    {
        int cur = first;
        double sc = liftedS<CP>(lp, cur, D);
        for (int i = 0; i < nv; ++i) {
            const int nxt = (cur + 1 == nv) ? 0 : cur + 1;
            const double sn = liftedS<CP>(lp, nxt, D);
            const d3 pc = lp.fp(cur);
            if (sc < 0) {
                if (np == 0) { P0 = pc; sum = pc; } else { sum += pc; if (np == 1) P1 = pc; else if (np == 2) P2 = pc; }
                np++;
            }
            if ((sc * sn) < 0) {
                const double w = sc / (sc - sn);
                const d3 cp = pc + w * (lp.fp(nxt) - pc);
                if (np == 0) { P0 = cp; sum = cp; } else { sum += cp; if (np == 1) P1 = cp; else if (np == 2) P2 = cp; }
                np++;
                if (nip < CP::MAXIP) ip[nip] = cp; else err |= SVERR_IFACE_POINTS;
                nip++;
            }
            cur = nxt;
            sc = sn;
        }
    }
    if (nip > CP::MAXIP) nip = CP::MAXIP;
    if (np < 3) {
        centre = fullC;
        area = fullA;
        return -1;
    }
    if (np == 3) {
        centre = (1.0 / 3.0) * (P0 + P1 + P2);
        area = 0.5 * cross(P1 - P0, P2 - P0);
        return 0;
    }
    // pass 2: edges (sp[i], sp[i+1]) in order, closing with (sp[np-1], sp[0])
    d3 fC = sum;
    fC /= double(np);
    d3 sumN = zero3(), sumAc = zero3();
    double sumA = 0.0;
    d3 prev = P0;
    {
        int cur = first, k = 0, kc = 0;
        double sc = liftedS<CP>(lp, cur, D);
        for (int i = 0; i < nv; ++i) {
            const int nxt = (cur + 1 == nv) ? 0 : cur + 1;
            const double sn = liftedS<CP>(lp, nxt, D);
            if (sc < 0) {
                const d3 pc = lp.fp(cur);
                if (k > 0) {
                    const d3 c = prev + pc + fC;
                    const d3 n = cross(pc - prev, fC - prev);
                    const double a = mag(n);
                    sumN += n;
                    sumA += a;
                    sumAc += a * c;
                }
                prev = pc;
                k++;
            }
            if ((sc * sn) < 0) {
                const d3 cp = ip[kc < CP::MAXIP ? kc : CP::MAXIP - 1];  // the point pass 1 stored (overflow is flagged in err)
                kc++;
                if (k > 0) {
                    const d3 c = prev + cp + fC;
                    const d3 n = cross(cp - prev, fC - prev);
                    const double a = mag(n);
                    sumN += n;
                    sumA += a;
                    sumAc += a * c;
                }
                prev = cp;
                k++;
            }
            cur = nxt;
            sc = sn;
        }
    }
    {
        const d3 c = prev + P0 + fC;
        const d3 n = cross(P0 - prev, fC - prev);
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
    return 0;
}

#define SV_G 8  // lanes per cell
#ifndef SV_PLIC_THREADS
#define SV_PLIC_THREADS 128
#define SV_PLIC_MINB 4
#endif

// CTA = cpb cells x 8 lanes.  Face phases (A, C, E): thread t works for cell t/8, lane t%8.
// Leader phases (staging, B, D, F): the first cpb threads, one per cell (thread t <-> cell t), so the
// sequential reductions of 32 cells run on 32 lanes of ONE warp instead of 1 lane in each of 8 warps
// (measured on the first version: 9 of 32 lanes active on average, 10.8k warp-instructions per
// evaluation; the leader phases dominated).
template <class CP>
__global__ void __launch_bounds__(SV_PLIC_THREADS, SV_PLIC_MINB) k_plic_group(MeshDev m, const int* mixedCells, Ctl* ctl, const double* __restrict__ alpha,
                                                    const double* iN, int split, int* cellStatus, double* iD, double* iC, double* iS)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    GCellShared<CP>* shAll = reinterpret_cast<GCellShared<CP>*>(smemRaw);
    const int cpb = blockDim.x / SV_G;  // cells per block
    LanePriv<CP> lp;
    lp.base = reinterpret_cast<double*>(smemRaw + size_t(cpb) * sizeof(GCellShared<CP>)) + threadIdx.x;
    lp.T = blockDim.x;
    const int gidF = threadIdx.x / SV_G, lane = threadIdx.x % SV_G;
    // per-face records of this thread's cell (only variants whose cells can have more than SV_G local faces)
    const int recDoubles = FaceRec::doubles(m.maxFV);
    double* faceCache = reinterpret_cast<double*>(smemRaw + size_t(cpb) * sizeof(GCellShared<CP>)) + size_t(blockDim.x) * LanePriv<CP>::DOUBLES +
                        size_t(gidF) * (size_t(m.maxLocalFaces) * recDoubles + (m.maxLocalFaces + 1) / 2);
    int* faceNv = reinterpret_cast<int*>(faceCache + size_t(m.maxLocalFaces) * recDoubles);
    const bool leader = (threadIdx.x < cpb);
    GCellShared<CP>& shF = shAll[gidF];                      // the cell this thread clips faces for
    GCellShared<CP>& shL = shAll[leader ? threadIdx.x : 0];  // the cell this thread leads (if leader)
    const int nMixed = ctl->nMixed;
    const bool splitB = split != 0;
    int err = 0;

    // Persistent CTAs + an atomic batch counter: batches take 3..6 evaluations, and with a static grid of
    // ceil(nMixed/cpb) CTAs the last, partially filled wave cost a whole wave (measured 906 CTAs on 296 slots).
    __shared__ int sBase;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sBase = atomicAdd(&ctl->plicNext, 1) * cpb;
        __syncthreads();
        const int base = sBase;
        if (base >= nMixed) break;
        // ---- leader role: stage the polyhedron (local face list, sorted vertex distances), init the search
        const int iL = base + threadIdx.x;
        const bool validL = leader && iL < nMixed;
        const int cellL = validL ? mixedCells[iL] : 0;
        const double alphaI = validL ? alpha[cellL] : 0.5;
        const d3 nL = validL ? ld3(iN, cellL) : mk3(1.0, 0.0, 0.0);
        const double Vcell = validL ? __ldg(m.V + cellL) : 1.0;
        double lowDistance = 0, upDistance = 0, lowAlpha = 0.0, upAlpha = 1.0;
        int lowLabel = 0, upLabel = 0;
        double midLabel = 0, aOneThird = 0, deltaDistance = 0, curD = 0;
        int phase = 0;  // 0 bracketing, 1 collapsed bracket, 2 one third, 3 two thirds, 4 final
        int outStatus = 0;
        if (leader) {
            int nl = 0, nP = 0;
            if (validL) {
                const int c0 = __ldg(m.cellOff + cellL), c1 = __ldg(m.cellOff + cellL + 1);
                for (int k = c0; k < c1; ++k) {
                    const int f = __ldg(m.cellFaces + k);
                    if (splitB && !(__ldg(m.flat + f) > (1.0 - SV_TSMALL))) {
                        const int nv = __ldg(m.faceOff + f + 1) - __ldg(m.faceOff + f);
                        for (int t = 0; t < nv; ++t) {
                            if (nl < CP::MAXCF) {
                                shL.lfFace[nl] = f;
                                shL.lfTri[nl] = (short)t;
                                nl++;
                            } else err |= SVERR_CELL_FACES;
                        }
                    } else {
                        if (nl < CP::MAXCF) {
                            shL.lfFace[nl] = f;
                            shL.lfTri[nl] = -1;
                            nl++;
                        } else err |= SVERR_CELL_FACES;
                    }
                }
                // vertex distances, descending (values only: ties are indistinguishable, cutCell.C:664-670)
                const int p0 = __ldg(m.cellPtOff + cellL), p1 = __ldg(m.cellPtOff + cellL + 1);
                for (int k = p0; k < p1; ++k) {
                    if (nP < CP::MAXCP) shL.vd[nP++] = -dot(nL, ld3(m.points, __ldg(m.cellPts + k)));
                    else err |= SVERR_CELL_POINTS;
                }
                if (splitB) {
                    for (int k = c0; k < c1; ++k) {
                        const int f = __ldg(m.cellFaces + k);
                        if (!(__ldg(m.flat + f) > (1.0 - SV_TSMALL))) {
                            if (nP < CP::MAXCP) shL.vd[nP++] = -dot(nL, ld3(m.Cf, f));
                            else err |= SVERR_CELL_POINTS;
                        }
                    }
                }
                sortDescending(shL.vd, nP);
            }
            shL.nLocal = nl;
            shL.nPts = nP;
            // search state (cutCell.C:682-719)
            bool done = !validL;
            if (validL) {
                upLabel = nP - 1;
                lowDistance = shL.vd[0];
                upDistance = shL.vd[nP - 1];
                if (mag(nL) < SV_TSMALL) {  // cutCell.C:625-628: D/C/S stay untouched
                    outStatus = int(sgn(0.5 - alphaI));
                    done = true;
                } else if ((upLabel - lowLabel) > 1) {
                    midLabel = round(0.5 * (upLabel + lowLabel));
                    curD = shL.vd[int(midLabel)];
                    phase = 0;
                } else if (fabs(lowDistance - upDistance) < SV_TSMALL) {
                    curD = 0.5 * (lowDistance + upDistance);
                    phase = 1;
                } else {
                    deltaDistance = (upDistance - lowDistance) / 3.0;
                    curD = lowDistance + deltaDistance;
                    phase = 2;
                }
                shL.D = curD;
            }
            shL.active = done ? 0 : 1;
        }
        __syncthreads();

        // ---- face role: plane-independent data of this lane's face (when one face per lane) -----------
        const int iF = base + gidF;
        const bool validF = iF < nMixed;
        const int cellF = validF ? mixedCells[iF] : 0;
        const d3 nF = validF ? ld3(iN, cellF) : mk3(1.0, 0.0, 0.0);
        const int nLocal = shF.nLocal;
        const bool cached = nLocal <= SV_G;
        d3 fullC = zero3(), fullA = zero3();
        int nvC = 0;
        if (validF && shF.active && cached && lane < nLocal) {
            d3 fpC[CP::MAXFV];
            nvC = loadLocalFace<CP>(m, cellF, shF.lfFace[lane], shF.lfTri[lane], splitB, fpC, err);
            for (int q = 0; q < nvC; ++q) {
                lp.setFp(q, fpC[q]);
                lp.pn(q) = dot(fpC[q], nF);
            }
            fullC = faceCentreOF(fpC, nvC);
            fullA = faceAreaNormalOF(fpC, nvC);
        }
        if (CP::MAXCF > SV_G && validF && shF.active && !cached) {
            for (int k = lane; k < nLocal; k += SV_G) {
                d3 fp[CP::MAXFV];
                const int nv = loadLocalFace<CP>(m, cellF, shF.lfFace[k], shF.lfTri[k], splitB, fp, err);
                FaceRec rec{faceCache + size_t(k) * recDoubles, m.maxFV};
                for (int q = 0; q < nv; ++q) {
                    rec.setFp(q, fp[q]);
                    rec.pn(q) = dot(fp[q], nF);
                }
                rec.setFull(faceCentreOF(fp, nv), faceAreaNormalOF(fp, nv));
                faceNv[k] = nv;
            }
        }

        // ---- evaluation loop: one calcSubCell (cutCell.C:343-542) per iteration ----------------------
        while (__syncthreads_or(leader ? shL.active : 0)) {
            // A. clip the faces of this lane
            if (shF.active) {
                const double D = shF.D;
                for (int k = lane; k < nLocal; k += SV_G) {
                    GFaceRes<CP>& r = shF.res[k];
                    d3 c, a;
                    int nip, st;
                    if (CP::MAXCF <= SV_G || cached) {   // (the hex variant has no other path)
                        st = clipFaceStream<CP>(lp, nvC, D, fullC, fullA, c, a, r.ip, nip, err);
                    } else {
                        const FaceRec rec{faceCache + size_t(k) * recDoubles, m.maxFV};
                        st = clipFaceStream<CP>(rec, faceNv[k], D, rec.fullC(), rec.fullA(), c, a, r.ip, nip, err);
                    }
                    r.st = st;
                    r.nip = nip;
                    r.c = c;
                    r.a = a;
                }
            }
            __syncthreads();
            // B. leaders: classification + interface polygon centre (cutCell.C:37-54)
            const bool act = leader && shL.active;
            const int nLoc = leader ? shL.nLocal : 0;
            bool fullySubmerged = true, fullyEmpty = true;
            int nSubmergedFaces = 0, nCut = 0;
            if (act) {
                d3 fC = zero3();
                int nEp = 0;
                for (int k = 0; k < nLoc; ++k) {
                    const int st = shL.res[k].st;
                    if (st == 0) {
                        fullySubmerged = false;
                        fullyEmpty = false;
                        nCut++;
                        const int nip = shL.res[k].nip;
                        for (int q = 0; q < nip; ++q) {
                            fC += shL.res[k].ip[q];
                            nEp++;
                        }
                    } else if (st == -1) {
                        fullyEmpty = false;
                        nSubmergedFaces++;
                        nCut++;
                    } else {
                        fullySubmerged = false;
                    }
                }
                if (nEp > 0) fC /= double(nEp);
                shL.fC = fC;
                shL.cutAny = (!fullySubmerged && !fullyEmpty) ? 1 : 0;
            }
            __syncthreads();
            // C. lanes: interface segments of their cut faces (cutCell.C:60-79, the per-segment part)
            if (shF.active && shF.cutAny) {
                const d3 fC = shF.fC;
                for (int k = lane; k < nLocal; k += SV_G) {
                    const GFaceRes<CP>& r = shF.res[k];
                    if (r.st != 0) continue;
                    for (int pi = 0; pi < r.nip - 1; ++pi) {
                        const d3 p0 = r.ip[pi], nx = r.ip[pi + 1];
                        const d3 c = p0 + nx + fC;
                        const d3 nn = cross(nx - p0, fC - p0);
                        const double a = mag(nn);
                        shF.seg[k].nn[pi] = nn;
                        shF.seg[k].a[pi] = a;
                        shF.seg[k].ac[pi] = a * c;
                    }
                }
            }
            __syncthreads();
            // D. leaders: ordered accumulation, interface centre/area, cEst (cutCell.C:56-99,110)
            int status = 0;
            double VOF = 0.0;
            bool needVolume = false;
            d3 iCl = zero3(), iSl = zero3();
            if (act) {
                if (shL.cutAny) {
                    d3 sumN = zero3(), sumAc = zero3();
                    double sumA = 0.0;
                    for (int k = 0; k < nLoc; ++k) {
                        if (shL.res[k].st != 0) continue;
                        for (int pi = 0; pi < shL.res[k].nip - 1; ++pi) {
                            const d3 nn = shL.seg[k].nn[pi];
                            sumN += sgn(dot(nn, sumN)) * nn;
                            sumA += shL.seg[k].a[pi];
                            sumAc += shL.seg[k].ac[pi];
                        }
                    }
                    if (sumA < SV_ROOTVSMALL) {
                        iCl = shL.fC;
                        iSl = zero3();
                    } else {
                        iCl = (1.0 / 3.0) * sumAc / sumA;
                        iSl = 0.5 * sumN;
                    }
                    if (dot(iSl, iCl - zero3()) < 0.0) iSl = iSl * (-1.0);  // vs the origin: SURVEY 8a' item 24
                    if (mag(iSl) < SV_TSMALL) {
                        if (nSubmergedFaces == 0) {
                            status = 1;
                            VOF = 0.0;
                        } else {
                            status = -1;
                            VOF = 1.0;
                        }
                    } else {
                        status = 0;
                        needVolume = true;
                        d3 cEst = zero3();
                        for (int k = 0; k < nLoc; ++k)
                            if (shL.res[k].st <= 0) cEst += shL.res[k].c;
                        cEst += iCl;
                        cEst /= double(nCut + 1);
                        shL.cEst = cEst;
                    }
                } else if (fullyEmpty) {
                    status = 1;
                    VOF = 0.0;
                } else {
                    status = -1;
                    VOF = 1.0;
                }
                shL.cutAny = needVolume ? 1 : 0;
            }
            __syncthreads();
            // E. lanes: pyramid volumes of their sub-faces (cutCell.C:116-123)
            if (shF.active && shF.cutAny) {
                const d3 cEst = shF.cEst;
                for (int k = lane; k < nLocal; k += SV_G) {
                    const GFaceRes<CP>& r = shF.res[k];
                    if (r.st <= 0) shF.pv[k] = dmax(fabs(dot(r.a, r.c - cEst)), SV_VSMALL);
                }
            }
            __syncthreads();
            // F. leaders: ordered volume sum, then advance the search (cutCell.C:691-799)
            if (act) {
                if (needVolume) {
                    double vol = 0.0;
                    for (int k = 0; k < nLoc; ++k)
                        if (shL.res[k].st <= 0) vol += shL.pv[k];
                    vol += dmax(fabs(dot(iSl, iCl - shL.cEst)), SV_VSMALL);
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
                            curD = shL.vd[int(midLabel)];
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
                    iD[cellL] = curD;
                    st3(iC, cellL, iCl);
                    st3(iS, cellL, iSl);
                    shL.active = 0;
                } else {
                    shL.D = curD;
                }
            }
            // (the __syncthreads_or of the loop condition is the barrier after F)
        }
        if (validL) cellStatus[iL] = outStatus;
        __syncthreads();
    }
    if (err) atomicOr(&ctl->err, err);
}

}  // namespace svof
