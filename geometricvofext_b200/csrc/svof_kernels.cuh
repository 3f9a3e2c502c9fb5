// svof_kernels.cuh -- the CUDA kernels of the SimPLIC step (sm_100a, FP64, no tensor cores:
// nothing on this path is a dense contraction).  Kernel map:
//
//   once per mesh   k_face_geom, k_cell_geom, k_flatness_tetbase        (K0)
//   reconstruct()   k_clear_prev, k_mixed_bits, k_count_bits, k_scan_blocks, k_write_mixed,
//                   k_mark_near, k_ls_normals | k_alpha_grad_normals (K2), k_plic_group<Caps> (K3, svof_plic_group.cuh)
//   advect()        k_un0_worklist (K5), k_face_flux<Caps> (K6), k_dense_update (K7, THE
//                   streaming kernel: the only pass over all cells/faces), k_near_update,
//                   k_bound_deps / k_bound_run / k_bound_apply (K8: bounding as a dependency-counted DAG),
//                   k_near_finalize, k_alpha_bc (K9)
//
// Every list is built on the device and every launch has a size known on the host
// (grid-stride over device-side counts), so a step needs no device->host round trip.
#pragma once
#include "svof_geom.cuh"

namespace svof {

#define SV_MAX_SWEEPS 32

// device-resident control block (one per handle)
struct Ctl {
    int nMixed;   // mixedCells_.size()
    int nNear2;   // |near2|
    int nWork;    // (cut cell, downwind face) work items
    int nPhiPulled;  // zero-copy host path: phi entries read from the caller's pinned buffer in the last svof_step_host
    int phiUnsafe;   // host path: an out-of-bounds cell owns a face outside the phi bitmap (an EMPTY cell overfilled, Courant > 1): the step is redone with the full phi
    int packUnsafe;  // a bounding correction landed on a face outside the phi bitmap (host path: alphaPhi then comes back in full)
    int nRdf;     // cells of the isoRDF zone (mixed cells + their point neighbours)
    int nUCells;  // cells whose U the interface-velocity interpolation reads (end-to-end path)
    int plicNext; // batch counter of the persistent plane-positioning kernel
    int epoch;    // advect() counter kept on the DEVICE (a captured CUDA graph replays correctly): bounding tags derive from it
    int nDeltaA, nDeltaF;  // changed alpha cells / alphaPhi faces of the last delta read-back
    int err;      // SVERR_* flags
    int nOob[2];  // out-of-bounds lists (double buffered between sweeps)
    int nPend[SV_MAX_SWEEPS + 1];    // (unused)
    int nAff[SV_MAX_SWEEPS + 1];     // cells touched by the corrections of sweep s
    int nearOob[SV_MAX_SWEEPS + 1];  // # near2 cells violating the limitFlux loop condition after s sweeps
    int pad_;
    unsigned long long minDense, maxDense;  // keys over the cells outside near2 (streaming kernel)
    unsigned long long minNear0, maxNear0;  // over near2 before bounding
    unsigned long long minNearF, maxNearF;  // over near2 after bounding (before snap/clip)
    unsigned long long dbg[8];              // SV_BOUND_STATS instrumentation
};

struct StepParams {
    double mixedTol, snapTol;
    int clip, nAlphaBounds, split;
    int geomD[3];
};

__device__ __forceinline__ bool bitTest(const unsigned int* bits, int i) { return (bits[i >> 5] >> (i & 31)) & 1u; }

#ifndef SV_VARIANT  // the capacity-independent kernels live in the main translation unit only
// ============================================================================ K0 ====
// primitiveMeshTools::faceCentresAndAreas (OF, recalled) -- thread per face
__global__ void k_face_geom(MeshDev m, double* Cf, double* Sf, double* magSf, int haveGeom)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.nFaces) return;
    if (!haveGeom) {
        const int o = m.faceOff[f], n = m.faceOff[f + 1] - o;
        d3 fc, fa;
        if (n == 3) {
            const d3 p0 = ld3(m.points, m.facePts[o]), p1 = ld3(m.points, m.facePts[o + 1]), p2 = ld3(m.points, m.facePts[o + 2]);
            fc = (1.0 / 3.0) * (p0 + p1 + p2);
            fa = 0.5 * cross(p1 - p0, p2 - p0);
        } else {
            d3 sumN = zero3(), sumAc = zero3();
            double sumA = 0.0;
            d3 fCentre = ld3(m.points, m.facePts[o]);
            for (int pi = 1; pi < n; ++pi) fCentre += ld3(m.points, m.facePts[o + pi]);
            fCentre /= double(n);
            for (int pi = 0; pi < n; ++pi) {
                const d3 nextPoint = ld3(m.points, m.facePts[o + ((pi == n - 1) ? 0 : pi + 1)]);
                const d3 thisPoint = ld3(m.points, m.facePts[o + pi]);
                const d3 c = thisPoint + nextPoint + fCentre;
                const d3 nn = cross(nextPoint - thisPoint, fCentre - thisPoint);
                const double a = mag(nn);
                sumN += nn;
                sumA += a;
                sumAc += a * c;
            }
            if (sumA < SV_ROOTVSMALL) {
                fc = fCentre;
                fa = zero3();
            } else {
                fc = (1.0 / 3.0) * sumAc / sumA;
                fa = 0.5 * sumN;
            }
        }
        st3(Cf, f, fc);
        st3(Sf, f, fa);
    }
    magSf[f] = mag(mk3(Sf[3 * (int64_t)f], Sf[3 * (int64_t)f + 1], Sf[3 * (int64_t)f + 2]));
}

// primitiveMeshTools::cellCentresAndVols (OF, recalled) -- thread per cell over the cells() row,
// which visits owned faces (ascending) then neighbour-side faces (ascending): OF's own order.
__global__ void k_cell_geom(MeshDev m, double* C, double* V)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.nCells) return;
    const int c0 = m.cellOff[c], c1 = m.cellOff[c + 1];
    d3 cEst = zero3();
    for (int k = c0; k < c1; ++k) cEst += ld3(m.Cf, m.cellFaces[k]);
    cEst /= double(c1 - c0);
    d3 cc = zero3();
    double vol = 0.0;
    for (int k = c0; k < c1; ++k) {
        const int f = m.cellFaces[k];
        const d3 fc = ld3(m.Cf, f), fa = ld3(m.Sf, f);
        const double pyr3Vol = (m.owner[f] == c) ? dot(fa, fc - cEst) : dot(fa, cEst - fc);
        const d3 pc = (3.0 / 4.0) * fc + (1.0 / 4.0) * cEst;
        cc += pyr3Vol * pc;
        vol += pyr3Vol;
    }
    if (fabs(vol) > SV_VSMALL) cc /= vol; else cc = cEst;
    st3(C, c, cc);
    V[c] = vol * (1.0 / 3.0);
}

// reconstruction::updateFaceFlatness (reconstruction.C:408-440) + polyMesh::tetBasePtIs
__device__ __forceinline__ double tetQuality(const d3& a, const d3& b, const d3& c, const d3& d)
{
    const double vol = (1.0 / 6.0) * dot(cross(b - a, c - a), d - a);
    const d3 ea = b - a, eb = c - a, ec = d - a;
    const double lambda = magSqr(ec) - dot(ea, ec);
    const double mu = magSqr(eb) - dot(ea, eb);
    const d3 ba = cross(eb, ea), ca = cross(ec, ea);
    const d3 num = lambda * ba - mu * ca;
    const double denom = dot(ec, ba);
    double R = SV_GREAT;
    if (fabs(denom) >= SV_ROOTVSMALL) R = mag(0.5 * (ea + num / denom));
    const double Rm = dmin(R, SV_GREAT);
    return vol / ((8.0 / (9.0 * sqrt(3.0))) * (Rm * (Rm * Rm)) + SV_ROOTVSMALL);
}

__global__ void k_flatness_tetbase(MeshDev m, double* flat, unsigned char* tetBase)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.nFaces) return;
    const int o = m.faceOff[f], n = m.faceOff[f + 1] - o;
    double fl = 1.0;
    if (n > 3 && m.magSf[f] > SV_ROOTVSMALL) {
        const d3 fc = ld3(m.Cf, f);
        double sumA = 0.0;
        for (int pi = 0; pi < n; ++pi) {
            const d3 thisPoint = ld3(m.points, m.facePts[o + pi]);
            const d3 nextPoint = ld3(m.points, m.facePts[o + ((pi + 1) % n)]);
            sumA += mag(0.5 * cross(nextPoint - thisPoint, fc - thisPoint));
        }
        fl = m.magSf[f] / (sumA + SV_ROOTVSMALL);
    }
    flat[f] = fl;
    // polyMeshTetDecomposition::findFaceBasePts, minTetQuality 1e-9
    const bool internal = f < m.nIF;
    const d3 oCc = ld3(m.C, m.owner[f]);
    const d3 nCc = internal ? ld3(m.C, m.neighbour[f]) : zero3();
    int found = -1;
    for (int base = 0; base < n && found < 0; ++base) {
        double minQ = SV_VGREAT;
        const d3 pb = ld3(m.points, m.facePts[o + base]);
        for (int t = 1; t < n - 1; ++t) {
            const int ia = (t + base) % n, ib = (ia + 1) % n;
            const d3 pa = ld3(m.points, m.facePts[o + ia]), pbb = ld3(m.points, m.facePts[o + ib]);
            double q = tetQuality(oCc, pb, pa, pbb);
            if (internal) q = dmin(q, tetQuality(nCc, pb, pbb, pa));
            if (q < minQ) minQ = q;
        }
        if (minQ > 1e-9) found = base;
    }
    tetBase[f] = (unsigned char)((found < 0) ? 0 : found);
}

// ==================================================================== reconstruct ====
// A1 (reconstruction.C:665-672, reconstruction.H:281-288): one bit per cell, one coalesced word per warp
__global__ void k_mixed_bits(const double* __restrict__ alpha, int nCells, double tol, unsigned int* bits)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    bool mixed = false;
    if (c < nCells) {
        const double a = alpha[c];
        mixed = (tol < a) && (a < 1.0 - tol);
    }
    const unsigned int w = __ballot_sync(0xffffffffu, mixed);
    if ((threadIdx.x & 31) == 0 && c < nCells) bits[c >> 5] = w;
}

// overset meshes (reconstruction.C:649-662): only CALCULATED cells may enter the interface-cell list; `calc` holds one bit
// per CALCULATED cell (svof_set_cell_types)
__global__ void k_and_bits(unsigned int* __restrict__ bits, const unsigned int* __restrict__ calc, int nWords)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nWords) bits[w] &= calc[w];
}

// alpha[idx[i]] = vals[i] (halo cells refreshed from their owning rank) with the mixed-cell bitmap kept up to date,
// so a decomposed run needs no dense bitmap pass after its halo swap
__global__ void k_scatter_alpha(const int* __restrict__ idx, const double* __restrict__ vals, long long n, double tol, double* alpha,
                                unsigned int* bits, int updateBits)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = idx[i];
    const double a = vals[i];
    alpha[c] = a;
    if (updateBits) {
        const unsigned int bit = 1u << (c & 31);
        if ((tol < a) && (a < 1.0 - tol)) atomicOr(&bits[c >> 5], bit);
        else atomicAnd(&bits[c >> 5], ~bit);
    }
}

// ---- the front of reconstruct(): three launches ------------------------------------------------------------------
//   k_front_count : popcount of the mixed-cell bitmap per 1024-word block  +  the sparse clears of the previous step
//                   (interfaceN/D/C/S and cellSlot of the previous mixed cells: the reference zero-fills 80 B/cell,
//                   reconstruction.C:636-641; the near1/near2 bitmap words of the previous near2 list instead of two
//                   field-sized memsets)
//   k_front_scan  : exclusive scan of the block sums (one CTA) + every per-step reset of the control block
//   k_front_write : ordered compaction (ascending list => mixedCells_ bit-exact in ORDER too), then each CTA marks the
//                   near sets of the list segment it just wrote
// (round 1: seven launches and two memsets)
#define SV_SCAN_WORDS 1024  // words per block
__global__ void __launch_bounds__(SV_SCAN_WORDS) k_front_count(const unsigned int* bits, int nWords, unsigned int* blockSums,
                                                               const int* mixedPrev, const int* near2Prev, int capNear, const Ctl* ctl,
                                                               double* iN, double* iD, double* iC, double* iS, int* cellSlot,
                                                               unsigned int* near1, unsigned int* near2)
{
    __shared__ unsigned int red[32];
    const int w = blockIdx.x * SV_SCAN_WORDS + threadIdx.x;
    unsigned int cnt = (w < nWords) ? __popc(bits[w]) : 0;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned int v = red[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) blockSums[blockIdx.x] = v;
    }
    // the previous reconstruct's lists (their counts are still in the control block: k_front_scan resets them)
    const int nM = ctl->nMixed, nN = min(ctl->nNear2, capNear);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nM; i += stride) {
        const int c = mixedPrev[i];
        st3(iN, c, zero3());
        st3(iC, c, zero3());
        st3(iS, c, zero3());
        iD[c] = 0.0;
        cellSlot[c] = -1;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nN; i += stride) {
        const int c = near2Prev[i];
        near1[c >> 5] = 0u;
        near2[c >> 5] = 0u;
    }
}

__global__ void k_front_scan(unsigned int* blockSums, int nBlocks, Ctl* ctl, int capacity)
{
    // single CTA, exclusive scan in place; total -> ctl->nMixed
    __shared__ unsigned int warpTot[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nBlocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned int v = (i < nBlocks) ? blockSums[i] : 0;
        unsigned int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned int t = warpTot[threadIdx.x], ti = t;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, ti, o);
                if (threadIdx.x >= o) ti += u;
            }
            warpTot[threadIdx.x] = ti - t;  // exclusive
        }
        __syncthreads();
        const unsigned int excl = carry + warpTot[threadIdx.x >> 5] + inc - v;
        if (i < nBlocks) blockSums[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int n = (int)carry;
        if (n > capacity) {
            atomicOr(&ctl->err, SVERR_LIST);
            n = capacity;
        }
        ctl->nMixed = n;
        // per-step resets: what reconstruct() and the advect() that follows it start from
        ctl->nNear2 = 0;
        ctl->plicNext = 0;
        ctl->minDense = ~0ull;
        ctl->maxDense = 0ull;
        ctl->nWork = 0;
        ctl->epoch++;
        ctl->nOob[0] = ctl->nOob[1] = 0;
        for (int s = 0; s <= SV_MAX_SWEEPS; ++s) ctl->nPend[s] = ctl->nAff[s] = ctl->nearOob[s] = 0;
        ctl->minNear0 = ctl->minNearF = ~0ull;
        ctl->maxNear0 = ctl->maxNearF = 0ull;
    }
}

// near1 = mixed U face-neighbours (== needBounding of advectionTemplates.C:138-141 via
// advection.C:54-82); near2 = near1 U face-neighbours (cells a bounding correction can touch).
// Bits are set with atomicOr; the thread that flips a near2 bit appends the cell to the list,
// so the list is a duplicate-free SET (its order is irrelevant: every consumer is per-cell).
// Most marks are duplicates (43 per mixed cell, ~3 of them new), and atomics on the same bitmap word serialise in L2:
// a plain L2 read filters the duplicates first (a stale read only costs a redundant atomic).  8 lanes per mixed cell,
// one face-neighbour (and its own neighbours) per lane.
__device__ __forceinline__ void markNear2(int c, unsigned int* near2, int* near2List, Ctl* ctl, int cap)
{
    const unsigned int bit = 1u << (c & 31);
    if (__ldcg(&near2[c >> 5]) & bit) return;
    const unsigned int old = atomicOr(&near2[c >> 5], bit);
    if (!(old & bit)) {
        const int pos = atomicAdd(&ctl->nNear2, 1);
        if (pos < cap) near2List[pos] = c; else atomicOr(&ctl->err, SVERR_LIST);
    }
}
__device__ __forceinline__ void markNear1(int c, unsigned int* near1)
{
    const unsigned int bit = 1u << (c & 31);
    if (!(__ldcg(&near1[c >> 5]) & bit)) atomicOr(&near1[c >> 5], bit);
}

__global__ void __launch_bounds__(SV_SCAN_WORDS) k_front_write(MeshDev m, const unsigned int* bits, int nWords, const unsigned int* blockSums,
                                                               int capacity, int* mixedCells, int* cellStatus, int* cellSlot, Ctl* ctl,
                                                               unsigned int* near1, unsigned int* near2, int* near2List, int capNear)
{
    __shared__ unsigned int warpTot[32];
    __shared__ unsigned int blockTotal;
    const int w = blockIdx.x * SV_SCAN_WORDS + threadIdx.x;
    const unsigned int word = (w < nWords) ? bits[w] : 0;
    const unsigned int v = __popc(word);
    unsigned int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned int t = warpTot[threadIdx.x], ti = t;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, ti, o);
            if (threadIdx.x >= o) ti += u;
        }
        warpTot[threadIdx.x] = ti - t;
        if (threadIdx.x == 31) blockTotal = ti;
    }
    __syncthreads();
    const unsigned int segBegin = blockSums[blockIdx.x];
    unsigned int pos = segBegin + warpTot[threadIdx.x >> 5] + inc - v;
    unsigned int rem = word;
    while (rem) {
        const int b = __ffs(rem) - 1;
        rem &= rem - 1;
        const int c = (w << 5) + b;
        if ((int)pos < capacity) {
            mixedCells[pos] = c;
            cellStatus[pos] = -100;  // reconstruction.C:670
            cellSlot[c] = (int)pos;
        }
        ++pos;
    }
    __syncthreads();
    // the near sets of this CTA's list segment, 8 lanes per mixed cell
    const int segEnd = min((int)(segBegin + blockTotal), capacity);
    const int lane = threadIdx.x & 7;
    for (int i = (int)segBegin + (threadIdx.x >> 3); i < segEnd; i += SV_SCAN_WORDS >> 3) {
        const int c = mixedCells[i];
        if (lane == 0) {
            markNear1(c, near1);
            markNear2(c, near2, near2List, ctl, capNear);
        }
        for (int k = m.cellOff[c] + lane; k < m.cellOff[c + 1]; k += 8) {
            const int y = m.cellAsc[k].y;
            if (y < 0) continue;
            markNear1(y, near1);
            markNear2(y, near2, near2List, ctl, capNear);
            for (int q = m.cellOff[y]; q < m.cellOff[y + 1]; ++q) {
                const int z = m.cellAsc[q].y;
                if (z >= 0) markNear2(z, near2, near2List, ctl, capNear);
            }
        }
    }
}

// A2: reconstruction::calcInterfaceNFromIsoAlphaGrad (reconstruction.C:85-141), thread per mixed
// cell.  Stencil = the cell, then every cell sharing a vertex (ascending label), then the valid
// boundary faces at its vertices (ascending): zoneCPCStencil membership with a pinned order.
// The dense normalisation pass of :138 collapses to the mixed cells (0/(0+SMALL) = 0 elsewhere).
#define SV_MAXST 160
#define SV_MAXSB 64
__device__ __forceinline__ bool sortedInsert(int* a, int& n, int cap, int v)
{
    int j = n - 1;
    while (j >= 0 && a[j] > v) --j;
    if (j >= 0 && a[j] == v) return true;
    if (n >= cap) return false;
    for (int q = n - 1; q > j; --q) a[q + 1] = a[q];
    a[j + 1] = v;
    ++n;
    return true;
}
// zoneCPCStencil of celli without the cell itself: cells sharing a vertex (ascending label) and valid boundary faces at its
// vertices (ascending), in thread-local lists
__device__ __forceinline__ void cpcStencilDev(const MeshDev& m, int celli, int* st, int& ns, int* sb, int& nb, int& err)
{
    ns = 0;
    nb = 0;
    for (int k = m.cellPtOff[celli]; k < m.cellPtOff[celli + 1]; ++k) {
        const int p = m.cellPts[k];
        for (int j = m.ptCellOff[p]; j < m.ptCellOff[p + 1]; ++j) {
            const int c = m.ptCells[j];
            if (c != celli && !sortedInsert(st, ns, SV_MAXST, c)) err |= SVERR_STENCIL;
        }
        for (int j = m.ptBFOff[p]; j < m.ptBFOff[p + 1]; ++j) {
            const int bf = m.ptBFaces[j];
            if (m.bKind[bf] == 0 && !sortedInsert(sb, nb, SV_MAXSB, bf)) err |= SVERR_STENCIL;
        }
    }
}
// leastSquareGrad<scalar>("polyDegree1", geometricD).grad over {celli} + stencil: cell values / boundary-face values
__device__ __forceinline__ d3 lsGradDev(const MeshDev& m, const StepParams& sp, int celli, const int* st, int ns, const int* sb, int nb,
                                        const double* __restrict__ cellVal, const double* __restrict__ bVal)
{
    int dims[3], nDims = 0;
    for (int d = 0; d < 3; ++d)
        if (sp.geomD[d] == 1) dims[nDims++] = d;
    const int nTerms = 1 + nDims;
    double A[4][4], src[4];
    for (int r = 0; r < 4; ++r) {
        src[r] = 0.0;
        for (int q = 0; q < 4; ++q) A[r][q] = 0.0;
    }
    const d3 Ci = ld3(m.C, celli);
    for (int s = -1; s < ns + nb; ++s) {
        d3 pos;
        double val;
        if (s < 0) {
            pos = Ci;
            val = cellVal[celli];
        } else if (s < ns) {
            pos = ld3(m.C, st[s]);
            val = cellVal[st[s]];
        } else {
            const int bf = sb[s - ns];
            pos = ld3(m.Cf, m.nIF + bf);
            val = bVal[bf];
        }
        pos -= Ci;
        const double comp[3] = {pos.x, pos.y, pos.z};
        double terms[4];
        terms[0] = 1.0;
        for (int d = 0; d < nDims; ++d) terms[d + 1] = comp[dims[d]];
        for (int r = 0; r < nTerms; ++r) {
            src[r] += terms[r] * val;
            for (int q = 0; q < nTerms; ++q) A[r][q] += terms[r] * terms[q];
        }
    }
    luSolve4(A, src, nTerms);
    double g[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < nDims; ++d) g[dims[d]] = src[d + 1];
    return mk3(g[0], g[1], g[2]);
}
__global__ void __launch_bounds__(128) k_ls_normals(MeshDev m, const int* mixedCells, Ctl* ctl, const double* __restrict__ alpha,
                                                    const double* __restrict__ alphaB, StepParams sp, double* iN)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int celli = mixedCells[i];
        int st[SV_MAXST], sb[SV_MAXSB];
        int ns = 0, nb = 0, err = 0;
        cpcStencilDev(m, celli, st, ns, sb, nb, err);
        if (err) atomicOr(&ctl->err, err);
        d3 nn = -lsGradDev(m, sp, celli, st, ns, sb, nb, alpha, alphaB);
        nn /= (mag(nn) + SV_SMALL);
        st3(iN, celli, nn);
    }
}

// ---- orientationMethod isoRDF (reconstruction::calcInterfaceNFromIsoRDF, reconstruction.C:196-405) ---------------------
// The iteration is driven from the host (svof_b200.cu:rdfNormals): its convergence test needs sums over the mixed cells
// in list order.  Not a hot path: no shipped case selects it; thread per item.

// reconstructedDistanceFunction::markCellsNearSurf(isMixedCell, 1) (OF, recalled): the mixed cells and every cell sharing
// a vertex with one, as bitmap + list
__global__ void k_rdf_mark(MeshDev m, const int* mixedCells, Ctl* ctl, unsigned int* bits, int* list, int cap)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = mixedCells[i];
        for (int k = m.cellPtOff[c]; k < m.cellPtOff[c + 1]; ++k) {
            const int p = m.cellPts[k];
            for (int j = m.ptCellOff[p]; j < m.ptCellOff[p + 1]; ++j) {   // the cell's own vertices list the cell itself
                const int y = m.ptCells[j];
                const unsigned int bit = 1u << (y & 31);
                if (bits[y >> 5] & bit) continue;
                const unsigned int old = atomicOr(&bits[y >> 5], bit);
                if (!(old & bit)) {
                    const int pos = atomicAdd(&ctl->nRdf, 1);
                    if (pos < cap) list[pos] = y; else atomicOr(&ctl->err, SVERR_LIST);
                }
            }
        }
    }
}
// a fresh RDF object per call (reconstruction.C:204): zero where this call will read; forget the previous call's marks
__global__ void k_rdf_reset(MeshDev m, const int* list, int n, unsigned int* bits, double* RDF, double* RDFb, int clearBits)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = list[i];
        if (clearBits) { bits[c >> 5] = 0u; continue; }
        RDF[c] = 0.0;
        for (int k = m.cellOff[c]; k < m.cellOff[c + 1]; ++k) {
            const int f = m.cellFaces[k];
            if (f >= m.nIF) RDFb[f - m.nIF] = 0.0;
        }
    }
}
// isoInterfaceGrad (reconstruction.C:144-192): interfaceNormal[i] = LS gradient of a field (alpha, then the RDF)
__global__ void __launch_bounds__(128) k_rdf_grad(MeshDev m, const int* mixedCells, Ctl* ctl, const double* __restrict__ cellVal,
                                                  const double* __restrict__ bVal, StepParams sp, double* iNormal)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int celli = mixedCells[i];
        int st[SV_MAXST], sb[SV_MAXSB];
        int ns = 0, nb = 0, err = 0;
        cpcStencilDev(m, celli, st, ns, sb, nb, err);
        if (err) atomicOr(&ctl->err, err);
        st3(iNormal, i, lsGradDev(m, sp, celli, st, ns, sb, nb, cellVal, bVal));
    }
}
// Vector::normalise(tol) (OF, recalled), in place
__device__ __forceinline__ d3 normaliseOF(d3 v, double tol)
{
    const double s = mag(v);
    if (s < tol) return zero3();
    v /= s;
    return v;
}
// reconstruction.C:234-235: interfaceN = -interfaceNormal[i].normalise(SMALL) (the normalisation stays in interfaceNormal);
// cells flagged too coarse keep D/C/S/status of their last evaluation: saved here, put back by k_rdf_restore after the
// plane-positioning kernel has run over ALL mixed cells
__global__ void k_rdf_set_normals(const int* mixedCells, Ctl* ctl, double* iNormal, double* iN, const unsigned char* coarse,
                                  const int* cellStatus, const double* iD, const double* iC, const double* iS, double* save)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = mixedCells[i];
        const d3 v = normaliseOF(ld3(iNormal, i), SV_SMALL);
        st3(iNormal, i, v);
        st3(iN, c, -v);
        if (coarse[i]) {
            double* s = save + 8 * (size_t)i;
            s[0] = iD[c];
            s[1] = iC[3 * (size_t)c]; s[2] = iC[3 * (size_t)c + 1]; s[3] = iC[3 * (size_t)c + 2];
            s[4] = iS[3 * (size_t)c]; s[5] = iS[3 * (size_t)c + 1]; s[6] = iS[3 * (size_t)c + 2];
            s[7] = (double)cellStatus[i];
        }
    }
}
__global__ void k_rdf_restore(const int* mixedCells, Ctl* ctl, const unsigned char* coarse, int* cellStatus, double* iD, double* iC,
                              double* iS, const double* save)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!coarse[i]) continue;
        const int c = mixedCells[i];
        const double* s = save + 8 * (size_t)i;
        iD[c] = s[0];
        iC[3 * (size_t)c] = s[1]; iC[3 * (size_t)c + 1] = s[2]; iC[3 * (size_t)c + 2] = s[3];
        iS[3 * (size_t)c] = s[4]; iS[3 * (size_t)c + 1] = s[5]; iS[3 * (size_t)c + 2] = s[6];
        cellStatus[i] = (int)s[7];
    }
}
// reconstructedDistanceFunction::constructRDF (OF v2312, recalled; reconstruction.C:257-264 with centre = interfaceC,
// normal = -interfaceN): weighted average of the distances to the planes of the stencil cells
__device__ __forceinline__ void rdfAverage(const MeshDev& m, const d3& p, int celli, const int* st, int ns, const double* iN,
                                           const double* iC, double& averageDist, double& avgWeight)
{
    averageDist = 0.0;
    avgWeight = 0.0;
    for (int s = -1; s < ns; ++s) {   // the cell itself first, then its point neighbours ascending; patch values of interfaceN are zero
        const int g = (s < 0) ? celli : st[s];
        d3 n = ld3(iN, g);
        if (mag(n) != 0.0) {
            n /= mag(n);
            d3 d = ld3(iC, g) - p;
            const double distToSurf = dot(d, n);
            double weight;
            if (mag(d) != 0.0) {
                d /= mag(d);
                const double a = fabs(dot(d, n));
                weight = a * a;
            } else {
                weight = 1.0;
            }
            averageDist += distToSurf * weight;
            avgWeight += weight;
        }
    }
}
__global__ void __launch_bounds__(128) k_rdf_construct(MeshDev m, const int* list, Ctl* ctl, const double* iN, const double* iC, double* RDF,
                                                       double* RDFb)
{
    const int n = ctl->nRdf;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int celli = list[i];
        int st[SV_MAXST], sb[SV_MAXSB];
        int ns = 0, nb = 0, err = 0;
        bool haveSt = false;
        const d3 nI = ld3(iN, celli);
        if (mag(nI) != 0.0) {   // interface cell: the distance of its own plane
            const d3 nn = nI / mag(nI);
            RDF[celli] = dot(ld3(iC, celli) - ld3(m.C, celli), nn);
        } else {
            cpcStencilDev(m, celli, st, ns, sb, nb, err);
            haveSt = true;
            double ad, aw;
            rdfAverage(m, ld3(m.C, celli), celli, st, ns, iN, iC, ad, aw);
            if (aw != 0.0) RDF[celli] = ad / aw;
        }
        for (int k = m.cellOff[celli]; k < m.cellOff[celli + 1]; ++k) {   // calculated patches: the same average at the face centre
            const int f = m.cellFaces[k];
            if (f < m.nIF || m.bKind[f - m.nIF] != 0) continue;
            if (!haveSt) {
                cpcStencilDev(m, celli, st, ns, sb, nb, err);
                haveSt = true;
            }
            double ad, aw;
            rdfAverage(m, ld3(m.Cf, f), celli, st, ns, iN, iC, ad, aw);
            RDFb[f - m.nIF] = (aw != 0.0) ? ad / aw : 0.0;
        }
        if (err) atomicOr(&ctl->err, err);
    }
}
// reconstruction.C:277-340: per mixed cell the normal residual and the mean angle to the normals of its stencil
__global__ void __launch_bounds__(128) k_rdf_residual(MeshDev m, const int* mixedCells, Ctl* ctl, const double* iN, double* iNormal,
                                                      double* normalRes, double* avgAngle)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int celli = mixedCells[i];
        int st[SV_MAXST], sb[SV_MAXSB];
        int ns = 0, nb = 0, err = 0;
        cpcStencilDev(m, celli, st, ns, sb, nb, err);
        if (err) atomicOr(&ctl->err, err);
        const d3 cellNormal = ld3(iN, celli);
        d3 v = ld3(iNormal, i);
        // `mag(N) < TSMALL or mag(interfaceNormal.normalise(SMALL)) < TSMALL`: the second operand (and its in-place
        // normalisation) is only evaluated when the first is false
        if (!(mag(cellNormal) < SV_TSMALL)) v = normaliseOF(v, SV_SMALL);
        double avgDiffNormal = 0.0, maxDiffNormal = SV_GREAT, weight = 0.0;
        for (int s = 0; s < ns; ++s) {   // j != 0: the cell itself is skipped; patch values of interfaceN are zero
            const d3 normal = ld3(iN, st[s]);
            const double mg = mag(normal);
            if (mg >= SV_TSMALL) {
                const d3 nn = normal / mg;
                const double cosAngle = dmax(dmin(dot(cellNormal, nn), 1.0), -1.0);
                avgDiffNormal += acos(cosAngle) * mg;
                weight += mg;
                if (cosAngle < maxDiffNormal) maxDiffNormal = cosAngle;
            }
        }
        if (weight != 0.0) avgDiffNormal /= weight; else avgDiffNormal = 0.0;
        v = normaliseOF(v, SV_SMALL);
        st3(iNormal, i, v);
        normalRes[i] = 1.0 - dot(cellNormal, -v);
        avgAngle[i] = avgDiffNormal;
    }
}

// A2, orientationMethod alphaGrad: reconstruction::calcInterfaceNFromRegAlphaGrad (reconstruction.C:74-82),
// volPointInterpolation::interpolate(alpha1) evaluated lazily at one point (OF, recalled; the scalar twin of pointU below)
__device__ double pointAlpha(const MeshDev& m, int p, const double* __restrict__ alpha, const double* __restrict__ alphaB)
{
    const d3 pt = ld3(m.points, p);
    double val = 0.0;
    if (!m.isPatchPoint[p]) {
        const int j0 = m.ptCellOff[p], j1 = m.ptCellOff[p + 1];
        double sumW = 0.0;
        for (int j = j0; j < j1; ++j) sumW += 1.0 / mag(pt - ld3(m.C, m.ptCells[j]));
        for (int j = j0; j < j1; ++j) {
            const int c = m.ptCells[j];
            const double pw = (1.0 / mag(pt - ld3(m.C, c))) / sumW;
            val += pw * alpha[c];
        }
        return val;
    }
    const int j0 = m.ptBFOff[p], j1 = m.ptBFOff[p + 1];
    double sumW = 0.0;
    for (int j = j0; j < j1; ++j) {
        const int bf = m.ptBFaces[j];
        if (m.bKind[bf] == 0) sumW += 1.0 / mag(pt - ld3(m.Cf, m.nIF + bf));
    }
    for (int j = j0; j < j1; ++j) {
        const int bf = m.ptBFaces[j];
        if (m.bKind[bf] != 0) continue;
        const double pw = (1.0 / mag(pt - ld3(m.Cf, m.nIF + bf))) / sumW;
        val += pw * alphaB[bf];
    }
    return val;
}

// surfaceInterpolation::weights() of internal face f (OF, recalled)
__device__ __forceinline__ double linearWeight(const MeshDev& m, int f)
{
    const d3 Sf = ld3(m.Sf, f), Cf = ld3(m.Cf, f);
    const double SfdOwn = fabs(dot(Sf, Cf - ld3(m.C, m.owner[f])));
    const double SfdNei = fabs(dot(Sf, ld3(m.C, m.neighbour[f]) - Cf));
    return (fabs(SfdOwn + SfdNei) > SV_ROOTVSMALL) ? SfdNei / (SfdOwn + SfdNei) : 0.5;
}

// pointLinear<scalar>::correction on internal face f (OF, recalled: pointLinear.C): the face value re-built from the
// point-interpolated field over the triangles (pi, f[k], f[k-1]) about pi = a C_P + (1 - a) C_N; lin = linearInterpolate(vf)[f].
// As published, a is mesh.weights() indexed by the OWNER CELL label; the quirk is kept (0.5 where that is no internal face).
__device__ double pointLinearCorrection(const MeshDev& m, int f, double lin, const double* __restrict__ alpha,
                                        const double* __restrict__ alphaB)
{
    const int P = m.owner[f], N = m.neighbour[f];
    const double a = (P < m.nIF) ? linearWeight(m, P) : 0.5;
    const d3 pi = a * ld3(m.C, P) + (1.0 - a) * ld3(m.C, N);
    const int q0 = m.faceOff[f], nv = m.faceOff[f + 1] - q0;
    const int pFirst = m.facePts[q0], pLast = m.facePts[q0 + nv - 1];
    double at = mag(0.5 * cross(ld3(m.points, pFirst) - pi, ld3(m.points, pLast) - pi));
    double sumAt = at;
    double sumPsip = at * (1.0 / 3.0) * (lin + pointAlpha(m, pFirst, alpha, alphaB) + pointAlpha(m, pLast, alpha, alphaB));
    int pPrev = pFirst;
    double vPrev = pointAlpha(m, pFirst, alpha, alphaB);
    for (int k = 1; k < nv; ++k) {
        const int pk = m.facePts[q0 + k];
        const double vk = pointAlpha(m, pk, alpha, alphaB);
        at = mag(0.5 * cross(ld3(m.points, pk) - pi, ld3(m.points, pPrev) - pi));
        sumAt += at;
        sumPsip += at * (1.0 / 3.0) * (lin + vk + vPrev);
        pPrev = pk;
        vPrev = vk;
    }
    return sumPsip / sumAt - lin;
}

// -fvc::grad(alpha1) with `Gauss linear` or `Gauss pointLinear` (OF, recalled: makeWeights, linear interpolate [+ the
// pointLinear correction], GaussGrad::calcGrad).
// Thread per mixed cell over its ascending-face row: the order GaussGrad accumulates in (internal faces ascending,
// then the patches).  Only mixed cells are evaluated (nothing on the path reads the others).
__global__ void __launch_bounds__(128) k_alpha_grad_normals(MeshDev m, const int* mixedCells, Ctl* ctl, const double* __restrict__ alpha,
                                                            const double* __restrict__ alphaB, double* iN, int pointLinear)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int celli = mixedCells[i];
        const d3 Ci = ld3(m.C, celli);
        const double ai = alpha[celli];
        d3 g = zero3();
        for (int k = m.cellOff[celli]; k < m.cellOff[celli + 1]; ++k) {
            const int2 e = m.cellAsc[k];
            const int f = e.x & 0x7fffffff;
            const bool isNei = e.x < 0;  // the cell is the neighbour of this face
            const d3 Sf = ld3(m.Sf, f);
            if (e.y >= 0) {
                const d3 Cf = ld3(m.Cf, f), Co = ld3(m.C, e.y);
                const double ao = alpha[e.y];
                const d3 CP = isNei ? Co : Ci, CN = isNei ? Ci : Co;
                const double aP = isNei ? ao : ai, aN = isNei ? ai : ao;
                const double SfdOwn = fabs(dot(Sf, Cf - CP));
                const double SfdNei = fabs(dot(Sf, CN - Cf));
                const double w = (fabs(SfdOwn + SfdNei) > SV_ROOTVSMALL) ? SfdNei / (SfdOwn + SfdNei) : 0.5;
                double af = w * (aP - aN) + aN;
                if (pointLinear) af += pointLinearCorrection(m, f, af, alpha, alphaB);
                const d3 t = Sf * af;
                if (isNei) g -= t;
                else g += t;
            } else {
                const int bf = -1 - e.y;
                if (m.bKind[bf] != 0) continue;  // empty (and processor) patches carry no value here
                g += Sf * alphaB[bf];
            }
        }
        g /= m.V[celli];
        d3 nn = -g;
        nn /= (mag(nn) + SV_SMALL);
        st3(iN, celli, nn);
    }
}

// ========================================================================= advect ====
// volPointInterpolation evaluated lazily at one point (OF, recalled)
__device__ d3 pointU(const MeshDev& m, int p, const double* __restrict__ U, const double* __restrict__ Ub)
{
    const d3 pt = ld3(m.points, p);
    d3 val = zero3();
    if (!m.isPatchPoint[p]) {
        const int j0 = m.ptCellOff[p], j1 = m.ptCellOff[p + 1];
        double sumW = 0.0;
        if (j1 - j0 <= 8) {
            // the inverse distances are needed twice (normalisation, then the weighted sum): keep them instead of
            // recomputing a square root and a divide per cell (a third of this kernel's instructions)
            double inv[8];
            int cs[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                inv[q] = 0.0;
                cs[q] = -1;
                if (j0 + q < j1) {
                    cs[q] = m.ptCells[j0 + q];
                    inv[q] = 1.0 / mag(pt - ld3(m.C, cs[q]));
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (cs[q] >= 0) sumW += inv[q];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (cs[q] < 0) continue;
                const double pw = inv[q] / sumW;
                val += pw * ld3(U, cs[q]);
            }
            return val;
        }
        for (int j = j0; j < j1; ++j) sumW += 1.0 / mag(pt - ld3(m.C, m.ptCells[j]));
        for (int j = j0; j < j1; ++j) {
            const int c = m.ptCells[j];
            const double pw = (1.0 / mag(pt - ld3(m.C, c))) / sumW;
            val += pw * ld3(U, c);
        }
        return val;
    }
    const int j0 = m.ptBFOff[p], j1 = m.ptBFOff[p + 1];
    double sumW = 0.0;
    for (int j = j0; j < j1; ++j) {
        const int bf = m.ptBFaces[j];
        if (m.bKind[bf] == 0) sumW += 1.0 / mag(pt - ld3(m.Cf, m.nIF + bf));
    }
    for (int j = j0; j < j1; ++j) {
        const int bf = m.ptBFaces[j];
        if (m.bKind[bf] != 0) continue;
        const double pw = (1.0 / mag(pt - ld3(m.Cf, m.nIF + bf))) / sumW;
        val += pw * ld3(Ub, bf);
    }
    return val;
}

// tetrahedron::pointToBarycentric (OF, recalled); a = cell centre
__device__ __forceinline__ double pointToBarycentric(const d3& a, const d3& b, const d3& c, const d3& d, const d3& pt, double* bary)
{
    const d3 v0 = a - d, v1 = b - d, v2 = c - d;
    const double xx = v0.x, xy = v1.x, xz = v2.x, yx = v0.y, yy = v1.y, yz = v2.y, zx = v0.z, zy = v1.z, zz = v2.z;
    const double detT = (xx * yy * zz + xy * yz * zx + xz * yx * zy - xx * yz * zy - xy * yx * zz - xz * yy * zx);
    if (fabs(detT) < SV_SMALL) {
        bary[0] = bary[1] = bary[2] = bary[3] = 0.25;
        return detT;
    }
    const double ixx = (yy * zz - zy * yz) / detT, ixy = (xz * zy - xy * zz) / detT, ixz = (xy * yz - xz * yy) / detT;
    const double iyx = (zx * yz - yx * zz) / detT, iyy = (xx * zz - xz * zx) / detT, iyz = (yx * xz - xx * yz) / detT;
    const double izx = (yx * zy - yy * zx) / detT, izy = (xy * zx - xx * zy) / detT, izz = (xx * yy - yx * xy) / detT;
    const d3 r = pt - d;
    const double rx = ixx * r.x + ixy * r.y + ixz * r.z;
    const double ry = iyx * r.x + iyy * r.y + iyz * r.z;
    const double rz = izx * r.x + izy * r.y + izz * r.z;
    bary[0] = rx;
    bary[1] = ry;
    bary[2] = rz;
    bary[3] = 1 - (rx + ry + rz);
    return detT;
}

__device__ __forceinline__ void tetTri(const MeshDev& m, int f, int tetPt, int celli, int* tri)
{
    const int o = m.faceOff[f], n = m.faceOff[f + 1] - o;
    const int base = m.tetBase[f];
    int facePtI = (tetPt + base) % n;
    int faceOtherPtI = (facePtI + 1) % n;
    if (m.owner[f] != celli) {
        const int t = facePtI;
        facePtI = faceOtherPtI;
        faceOtherPtI = t;
    }
    tri[0] = m.facePts[o + base];
    tri[1] = m.facePts[o + facePtI];
    tri[2] = m.facePts[o + faceOtherPtI];
}

// interpolationCellPoint<vector>::interpolate(position, celli) (OF, recalled)
__device__ d3 interpolateU(const MeshDev& m, const d3& position, int celli, const double* __restrict__ U,
                           const double* __restrict__ Ub)
{
    const double tol = SV_SMALL;
    const double cellVolume = m.V[celli];
    const d3 cc = ld3(m.C, celli);
    double w[4];
    int tri[3];
    bool found = false;
    const int c0 = m.cellOff[celli], c1 = m.cellOff[celli + 1];
    for (int k = c0; k < c1 && !found; ++k) {
        const int f = m.cellFaces[k];
        const int nv = m.faceOff[f + 1] - m.faceOff[f];
        for (int tetPt = 1; tetPt < nv - 1 && !found; ++tetPt) {
            tetTri(m, f, tetPt, celli, tri);
            const double det = pointToBarycentric(cc, ld3(m.points, tri[0]), ld3(m.points, tri[1]), ld3(m.points, tri[2]), position, w);
            if (fabs(det / cellVolume) > tol) {
                const double u = w[0], v = w[1], ww = w[2];
                if ((u + tol > 0) && (v + tol > 0) && (ww + tol > 0) && (u + v + ww < 1 + tol)) found = true;
            }
        }
    }
    if (!found) {  // least-violated tet (the interface centre lies inside its cell; safety net)
        double best = SV_VGREAT;
        int bt[3] = {0, 0, 0};
        double bw[4] = {0.25, 0.25, 0.25, 0.25};
        for (int k = c0; k < c1; ++k) {
            const int f = m.cellFaces[k];
            const int nv = m.faceOff[f + 1] - m.faceOff[f];
            for (int tetPt = 1; tetPt < nv - 1; ++tetPt) {
                int t3[3];
                double tw[4];
                tetTri(m, f, tetPt, celli, t3);
                pointToBarycentric(cc, ld3(m.points, t3[0]), ld3(m.points, t3[1]), ld3(m.points, t3[2]), position, tw);
                double viol = 0;
                for (int q = 0; q < 4; ++q) viol += (tw[q] < 0) ? -tw[q] : 0;
                if (viol < best) {
                    best = viol;
                    for (int q = 0; q < 3; ++q) bt[q] = t3[q];
                    for (int q = 0; q < 4; ++q) bw[q] = tw[q];
                }
            }
        }
        for (int q = 0; q < 3; ++q) tri[q] = bt[q];
        for (int q = 0; q < 4; ++q) w[q] = bw[q];
    }
    d3 t = ld3(U, celli) * w[0];
    t += pointU(m, tri[0], U, Ub) * w[1];
    t += pointU(m, tri[1], U, Ub) * w[2];
    t += pointU(m, tri[2], U, Ub) * w[3];
    return t;
}

// (Round 1 also tried 8 lanes per cut cell -- tets of different faces tested at once, the inverse-distance sums taken
// with shuffles: 87 us against 80 us for this thread-per-cell form.  The kernel is bound by its instruction count at
// 4 warps per scheduler (122 registers), not by the length of one thread's chain, and the shuffles added a third more.)
// A7 first half (advection.C:112-175): interface speed per cut cell + the compacted work list of
// (cut cell, downwind face) pairs.  Each face has exactly one upwind cell, so the list is
// duplicate free and the flux kernel's writes are conflict free.
__global__ void __launch_bounds__(128) k_un0_worklist(MeshDev m, const int* mixedCells, const int* cellStatus, Ctl* ctl,
                                                      const double* iN, const double* iC, const double* __restrict__ U,
                                                      const double* __restrict__ Ub, const double* __restrict__ phi,
                                                      double* Un0, int2* work, int capWork)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (cellStatus[i] != 0) {
            Un0[i] = 0.0;
            continue;
        }
        const int c = mixedCells[i];
        const d3 nn = ld3(iN, c);
        Un0[i] = dot(interpolateU(m, ld3(iC, c), c, U, Ub), nn);
        int cnt = 0;
        int loc[64];
        for (int k = m.cellOff[c]; k < m.cellOff[c + 1]; ++k) {
            const int f = m.cellFaces[k];
            bool down;
            if (f < m.nIF) down = (m.owner[f] == c) ? (phi[f] >= 0.0) : (phi[f] < 0.0);
            else down = (m.bKind[f - m.nIF] != 1) && (phi[f] >= 0.0);  // advection.C:192-197
            if (down) {
                if (cnt < 64) loc[cnt++] = f; else atomicOr(&ctl->err, SVERR_CELL_FACES);
            }
        }
        if (cnt) {
            const int pos = atomicAdd(&ctl->nWork, cnt);
            for (int q = 0; q < cnt; ++q) {
                if (pos + q < capWork) work[pos + q] = make_int2(i, loc[q]); else atomicOr(&ctl->err, SVERR_LIST);
            }
        }
    }
}

// A7 first half (advection.C:112-175) with 8 lanes per cut cell: interface speed Un0 = U_interp(interfaceC) . n and the
// compacted work list of (cut cell, downwind face) pairs.  interpolationCellPoint::interpolate (OF, recalled) walks the
// cell's tets in order (cells() face order, tet index ascending) until one contains the point: here lane g tests the
// tets of faces g, g+8, ... and the group takes the first hit in that order (ballot, lowest lane of the first round
// with a hit); the inverse-distance point values of the three tet vertices are formed with one (point, cell) pair per
// lane and the ORDERED sums redone by every lane from shared memory -- same operands, same order, same bits as the
// thread-per-cell form, a quarter of its dependent chain.  (Round 1's lane-cooperative attempt passed everything through
// shuffles and was instruction bound; this one stages through shared memory.)  Each face has exactly one upwind cell,
// so the work list is duplicate free and the flux kernel's writes are conflict free.
#define SV_UG_CELLS 16   // cut cells per 128-thread CTA
struct Un0Shared {
    double inv[8];
    double term[8][3];
};
__device__ __forceinline__ d3 pointUGroup(const MeshDev& m, int p, const double* __restrict__ U, const double* __restrict__ Ub, int g,
                                          unsigned gmask, int grpLane0, Un0Shared& sh)
{
    if (m.isPatchPoint[p] || (m.ptCellOff[p + 1] - m.ptCellOff[p]) > 8) {   // boundary points / many cells: one lane, in order
        d3 v = zero3();
        if (g == 0) v = pointU(m, p, U, Ub);
        v.x = __shfl_sync(gmask, v.x, grpLane0);
        v.y = __shfl_sync(gmask, v.y, grpLane0);
        v.z = __shfl_sync(gmask, v.z, grpLane0);
        return v;
    }
    const int j0 = m.ptCellOff[p], n = m.ptCellOff[p + 1] - j0;
    const d3 pt = ld3(m.points, p);
    int c = -1;
    double inv = 0.0;
    if (g < n) {
        c = m.ptCells[j0 + g];
        inv = 1.0 / mag(pt - ld3(m.C, c));
        sh.inv[g] = inv;
    }
    __syncwarp(gmask);
    double sumW = 0.0;
    for (int q = 0; q < n; ++q) sumW += sh.inv[q];
    if (g < n) {
        const double pw = inv / sumW;
        const d3 t = pw * ld3(U, c);
        sh.term[g][0] = t.x;
        sh.term[g][1] = t.y;
        sh.term[g][2] = t.z;
    }
    __syncwarp(gmask);
    d3 val = zero3();
    for (int q = 0; q < n; ++q) val += mk3(sh.term[q][0], sh.term[q][1], sh.term[q][2]);
    __syncwarp(gmask);   // the slots are reused by the next point
    return val;
}

__global__ void __launch_bounds__(128) k_un0_group(MeshDev m, const int* mixedCells, const int* cellStatus, Ctl* ctl, const double* iN,
                                                   const double* iC, const double* __restrict__ U, const double* __restrict__ Ub,
                                                   const double* __restrict__ phi, double* Un0, int2* work, int capWork)
{
    __shared__ Un0Shared shAll[SV_UG_CELLS];
    const int n = ctl->nMixed;
    const int lane = threadIdx.x & 31, g = lane & 7, grp = lane >> 3;
    const unsigned gmask = 0xFFu << (grp * 8);
    const int grpLane0 = grp * 8;
    Un0Shared& sh = shAll[threadIdx.x >> 3];
    const int groupsTotal = (gridDim.x * blockDim.x) >> 3;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < n; i += groupsTotal) {
        if (cellStatus[i] != 0) {
            if (g == 0) Un0[i] = 0.0;
            continue;
        }
        const int celli = mixedCells[i];
        const d3 position = ld3(iC, celli);
        const double tol = SV_SMALL;
        const double cellVolume = m.V[celli];
        const d3 cc = ld3(m.C, celli);
        const int c0 = m.cellOff[celli], c1 = m.cellOff[celli + 1];
        // ---- the tet that contains the interface centre: first hit in (face, tet) order
        double w[4] = {0.25, 0.25, 0.25, 0.25};
        int tri[3] = {0, 0, 0};
        bool found = false;
        for (int k0 = c0; k0 < c1 && !found; k0 += 8) {
            bool hit = false;
            const int k = k0 + g;
            if (k < c1) {
                const int f = m.cellFaces[k];
                const int nv = m.faceOff[f + 1] - m.faceOff[f];
                for (int tetPt = 1; tetPt < nv - 1 && !hit; ++tetPt) {
                    tetTri(m, f, tetPt, celli, tri);
                    const double det = pointToBarycentric(cc, ld3(m.points, tri[0]), ld3(m.points, tri[1]), ld3(m.points, tri[2]), position, w);
                    if (fabs(det / cellVolume) > tol) {
                        const double u = w[0], v = w[1], ww = w[2];
                        if ((u + tol > 0) && (v + tol > 0) && (ww + tol > 0) && (u + v + ww < 1 + tol)) hit = true;
                    }
                }
            }
            const unsigned b = __ballot_sync(gmask, hit) & gmask;
            if (b) {
                const int src = __ffs(b) - 1;
#pragma unroll
                for (int q = 0; q < 3; ++q) tri[q] = __shfl_sync(gmask, tri[q], src);
#pragma unroll
                for (int q = 0; q < 4; ++q) w[q] = __shfl_sync(gmask, w[q], src);
                found = true;
            }
        }
        if (!found) {  // least-violated tet (the interface centre lies inside its cell; safety net): one lane, in order
            if (g == 0) {
                double best = SV_VGREAT;
                for (int q = 0; q < 4; ++q) w[q] = 0.25;
                tri[0] = tri[1] = tri[2] = 0;
                for (int k = c0; k < c1; ++k) {
                    const int f = m.cellFaces[k];
                    const int nv = m.faceOff[f + 1] - m.faceOff[f];
                    for (int tetPt = 1; tetPt < nv - 1; ++tetPt) {
                        int t3[3];
                        double tw[4];
                        tetTri(m, f, tetPt, celli, t3);
                        pointToBarycentric(cc, ld3(m.points, t3[0]), ld3(m.points, t3[1]), ld3(m.points, t3[2]), position, tw);
                        double viol = 0;
                        for (int q = 0; q < 4; ++q) viol += (tw[q] < 0) ? -tw[q] : 0;
                        if (viol < best) {
                            best = viol;
                            for (int q = 0; q < 3; ++q) tri[q] = t3[q];
                            for (int q = 0; q < 4; ++q) w[q] = tw[q];
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) tri[q] = __shfl_sync(gmask, tri[q], grpLane0);
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = __shfl_sync(gmask, w[q], grpLane0);
        }
        // ---- cell value + the three point values, in the reference's order
        d3 t = ld3(U, celli) * w[0];
        t += pointUGroup(m, tri[0], U, Ub, g, gmask, grpLane0, sh) * w[1];
        t += pointUGroup(m, tri[1], U, Ub, g, gmask, grpLane0, sh) * w[2];
        t += pointUGroup(m, tri[2], U, Ub, g, gmask, grpLane0, sh) * w[3];
        if (g == 0) Un0[i] = dot(t, ld3(iN, celli));
        // ---- downwind faces of this cut cell -> work list (advection.C:134-151,192-197)
        for (int k0 = c0; k0 < c1; k0 += 8) {
            const int k = k0 + g;
            bool down = false;
            int f = 0;
            if (k < c1) {
                f = m.cellFaces[k];
                if (f < m.nIF) down = (m.owner[f] == celli) ? (phi[f] >= 0.0) : (phi[f] < 0.0);
                else down = (m.bKind[f - m.nIF] != 1) && (phi[f] >= 0.0);
            }
            const unsigned b = __ballot_sync(gmask, down) & gmask;
            if (b) {
                int pos = 0;
                if (g == 0) pos = atomicAdd(&ctl->nWork, __popc(b));
                pos = __shfl_sync(gmask, pos, grpLane0);
                if (down) {
                    const int slot = pos + __popc(b & ((1u << lane) - 1u));
                    if (slot < capWork) work[slot] = make_int2(i, f); else atomicOr(&ctl->err, SVERR_LIST);
                }
            }
        }
    }
}

__device__ __forceinline__ void blockMinMax(double mn, double mx, unsigned long long* gmin, unsigned long long* gmax)
{
    __shared__ unsigned long long smn[32], smx[32];
    unsigned long long kmn = dkey(mn), kmx = dkey(mx);
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_down_sync(0xffffffffu, kmn, o);
        const unsigned long long b = __shfl_down_sync(0xffffffffu, kmx, o);
        kmn = (a < kmn) ? a : kmn;
        kmx = (b > kmx) ? b : kmx;
    }
    if ((threadIdx.x & 31) == 0) {
        smn[threadIdx.x >> 5] = kmn;
        smx[threadIdx.x >> 5] = kmx;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        kmn = (threadIdx.x < nw) ? smn[threadIdx.x] : ~0ull;
        kmx = (threadIdx.x < nw) ? smx[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_down_sync(0xffffffffu, kmn, o);
            const unsigned long long b = __shfl_down_sync(0xffffffffu, kmx, o);
            kmn = (a < kmn) ? a : kmn;
            kmx = (b > kmx) ? b : kmx;
        }
        if (threadIdx.x == 0) {
            atomicMin(gmin, kmn);
            atomicMax(gmax, kmx);
        }
    }
}

// advection.C:291-308, applied per cell
__device__ __forceinline__ double snapClip(double a, double snapTol, int clip)
{
    if (snapTol > 0.0) a = a * pos0(a - snapTol) * neg0(a - (1.0 - snapTol)) + pos0(a - (1.0 - snapTol));
    if (clip) a = dmin(1.0, dmax(0.0, a));
    return a;
}

// upwind face transport of one face seen from cell c (advectionTemplates.C:371), boundary incl.
__device__ __forceinline__ bool upwindDVf(const MeshDev& m, int c, int f, bool flip, int other, double ph,
                                          const double* __restrict__ aOld, const double* __restrict__ alphaB, double dt, double& dvf)
{
    if (other >= 0) {
        const int own = flip ? other : c, nei = flip ? c : other;
        const double aUp = (ph >= 0) ? aOld[own] : aOld[nei];
        dvf = (ph * aUp) * dt;
        return true;
    }
    const int bf = -1 - other;
    const unsigned char kind = m.bKind[bf];
    if (kind == 1) return false;  // empty patch: no field
    double ab = alphaB[bf];
    if (kind == 2) ab = (ph >= 0) ? aOld[c] : ab;  // processor: upwind between the two sides
    dvf = (ph * ab) * dt;
    return true;
}

// exact-division shortcuts: x/y == x bitwise when x is +-0 (y finite, non-zero); most of a VOF
// domain is exactly empty, so the FP64 divides (~30 SASS instructions each) are skipped there
// The branch is made WARP-UNIFORM with a vote: nvcc if-converts a plain `x == 0 ? x : x / y` into an
// unconditional divide + select (measured: 5 divide subroutine calls per warp, 40% of all instructions
// of the streaming kernel, in a domain that is 98.6% empty).
__device__ __forceinline__ double divz(double x, double y)
{
    if (__any_sync(__activemask(), x != 0.0)) x = (x == 0.0) ? x : x / y;
    return x;
}

// warp min/max of doubles through their order-preserving keys with 4 REDUX instead of 20 SHFL
__device__ __forceinline__ unsigned long long warpMaxKey(unsigned long long k)
{
    const unsigned int hi = (unsigned int)(k >> 32);
    const unsigned int mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned int lo = (hi == mhi) ? (unsigned int)k : 0u;
    const unsigned int mlo = __reduce_max_sync(0xffffffffu, lo);
    return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ unsigned long long warpMinKey(unsigned long long k) { return ~warpMaxKey(~k); }

__device__ __forceinline__ void blockMinMaxFast(double mn, double mx, unsigned long long* gmin, unsigned long long* gmax)
{
    __shared__ unsigned long long smn[32], smx[32];
    const unsigned long long kmn = warpMinKey(dkey(mn)), kmx = warpMaxKey(dkey(mx));
    if ((threadIdx.x & 31) == 0) {
        smn[threadIdx.x >> 5] = kmn;
        smx[threadIdx.x >> 5] = kmx;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        const unsigned long long a = warpMinKey((threadIdx.x < nw) ? smn[threadIdx.x] : ~0ull);
        const unsigned long long b = warpMaxKey((threadIdx.x < nw) ? smx[threadIdx.x] : 0ull);
        if (threadIdx.x == 0) {
            atomicMin(gmin, a);
            atomicMax(gmax, b);
        }
    }
}

// A6 + A10 + A12 fused: THE streaming pass (K7).  Thread per cell; a cell gathers its faces in
// ascending face order through the cell->face CSR (deterministic segmented reduction, no atomics,
// same summation order as fvc::surfaceIntegrate), recomputes the upwind transport of each face on
// the fly (dVf is never materialised), writes alpha_new and, for the faces it owns, alphaPhi.
// Cells in near2 are left to the sparse kernels.  Also emits the next step's mixed-cell bitmap.
// (A sliced-ELL row layout with batched loads was tried in round 1: it coalesces the row loads but
// loses the L1 reuse of the 48-byte rows and needs 71 registers; measured 1.05 ms vs 0.58 ms.)
#ifndef SV_DENSE_UNROLL
#define SV_DENSE_UNROLL 1
#endif
constexpr int kDenseUnroll = SV_DENSE_UNROLL;
// L2 eviction hints for the streaming pass (PTX createpolicy + .L2::cache_hint), "dense_l2" option: the lines the pass
// streams (1.9 GB per launch through a 126 MB L2) become PREFERRED VICTIMS -- 1: connectivity rows, offsets, volumes and
// the stores; 2: those loads only; 3: everything incl. phi and alpha (measured: 3 costs the pass its own L2 reuse of
// phi / alpha, +0.09 ms) -- so the few MB the interface kernels on the other stream keep re-reading (mesh rows and fields next to the
// interface, the lists one kernel hands to the next) stay resident while it runs beside them.
template <bool EF> struct DenseMem;
template <> struct DenseMem<false> {
    __device__ __forceinline__ DenseMem() {}
    __device__ __forceinline__ int ldi(const int* p) const { return __ldg(p); }
    __device__ __forceinline__ int2 ldi2(const int2* p) const { return __ldg(p); }
    __device__ __forceinline__ double ldd(const double* p) const { return __ldg(p); }
    __device__ __forceinline__ void std_(double* p, double v) const { *p = v; }
};
template <> struct DenseMem<true> {
    unsigned long long pol;
    __device__ __forceinline__ DenseMem() { asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); }
    __device__ __forceinline__ int ldi(const int* p) const
    {
        int v;
        asm("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
        return v;
    }
    __device__ __forceinline__ int2 ldi2(const int2* p) const
    {
        int2 v;
        asm("ld.global.nc.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
        return v;
    }
    __device__ __forceinline__ double ldd(const double* p) const
    {
        double v;
        asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
        return v;
    }
    __device__ __forceinline__ void std_(double* p, double v) const
    {
        asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
    }
};

// the per-cell body of the streaming pass: returns whether the new value is a mixed cell; mn / mx take the unclipped value
template <int EF>
__device__ __forceinline__ bool denseCell(const MeshDev& m, int c, const double* __restrict__ aOld, double* __restrict__ aNew,
                                          const double* __restrict__ phi, const double* __restrict__ alphaB,
                                          double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2, double dt, double rDt,
                                          const double* __restrict__ Sp, const double* __restrict__ Su, const StepParams& sp, double& mn,
                                          double& mx)
{
    const DenseMem<(EF != 0)> mem;          // single-use streams: connectivity rows, offsets, volumes
    const DenseMem<(EF == 3)> memR;         // streams with reuse in L2 (phi: read by both cells of a face; alpha: gathers)
    const DenseMem<(EF == 1 || EF == 3)> memW;  // stores
    bool mixed = false;
    if (c < m.nCells && !bitTest(near2, c)) {
        const int k0 = mem.ldi(m.cellOff + c), k1 = mem.ldi(m.cellOff + c + 1);
        const double aC = memR.ldd(aOld + c);
        double sum = 0.0;
#pragma unroll kDenseUnroll
        for (int k = k0; k < k1; ++k) {
            const int2 e = mem.ldi2(m.cellAsc + k);
            const int f = e.x & 0x7fffffff;
            const bool flip = e.x < 0;
            const double ph = memR.ldd(phi + f);
            double aUp;
            if (e.y >= 0) {
                // upwind cell: owner if phi >= 0 else neighbour; this cell's own value is in a register
                const bool selfUp = flip ? (ph < 0) : (ph >= 0);
                aUp = selfUp ? aC : memR.ldd(aOld + e.y);
            } else {
                const int bf = -1 - e.y;
                const unsigned char kind = __ldg(m.bKind + bf);
                if (kind == 1) continue;  // empty patch: no field
                aUp = __ldg(alphaB + bf);
            }
            const double dvf = (ph * aUp) * dt;
            if (!flip) {
                sum += dvf;
                memW.std_(alphaPhi + f, divz(dvf, dt));
            } else {
                sum -= dvf;
            }
        }
        const double ivf = divz(sum, mem.ldd(m.V + c));
        double num = aC * rDt;
        if (Su) num = num + Su[c];
        num = num - ivf * rDt;
        double a = divz(num, (Sp ? (rDt - Sp[c]) : rDt));
        mn = dmin(mn, a);
        mx = dmax(mx, a);
        a = snapClip(a, sp.snapTol, sp.clip);
        memW.std_(aNew + c, a);
        mixed = (sp.mixedTol < a) && (a < 1.0 - sp.mixedTol);
    }
    return mixed;
}

template <int EF>
__device__ __forceinline__ void denseTile(const MeshDev& m, const double* __restrict__ aOld, double* __restrict__ aNew,
                                          const double* __restrict__ phi, const double* __restrict__ alphaB,
                                          double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2,
                                          unsigned int* __restrict__ mixedNext, double dt, double rDt, const double* __restrict__ Sp,
                                          const double* __restrict__ Su, const StepParams& sp, Ctl* ctl)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    double mn = SV_VGREAT, mx = -SV_VGREAT;
    const bool mixed = denseCell<EF>(m, c, aOld, aNew, phi, alphaB, alphaPhi, near2, dt, rDt, Sp, Su, sp, mn, mx);
    const unsigned int w = __ballot_sync(0xffffffffu, mixed);
    if ((threadIdx.x & 31) == 0 && c < m.nCells) mixedNext[c >> 5] = w;
    blockMinMaxFast(mn, mx, &ctl->minDense, &ctl->maxDense);
}
__global__ void __launch_bounds__(256, 8) k_dense_update(MeshDev m, const double* __restrict__ aOld, double* __restrict__ aNew,
                                                      const double* __restrict__ phi, const double* __restrict__ alphaB,
                                                      double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2,
                                                      unsigned int* __restrict__ mixedNext, double dt, double rDt,
                                                      const double* __restrict__ Sp, const double* __restrict__ Su, StepParams sp,
                                                      Ctl* ctl)
{
    denseTile<0>(m, aOld, aNew, phi, alphaB, alphaPhi, near2, mixedNext, dt, rDt, Sp, Su, sp, ctl);
}
// "dense_l2" variants (opt-in; all measured slower, profiles/r4f_sweep_dense_l2_hints.txt)
template <int EF>
__global__ void __launch_bounds__(256, 8) k_dense_update_hint(MeshDev m, const double* __restrict__ aOld, double* __restrict__ aNew,
                                                           const double* __restrict__ phi, const double* __restrict__ alphaB,
                                                           double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2,
                                                           unsigned int* __restrict__ mixedNext, double dt, double rDt,
                                                           const double* __restrict__ Sp, const double* __restrict__ Su, StepParams sp,
                                                           Ctl* ctl)
{
    denseTile<EF>(m, aOld, aNew, phi, alphaB, alphaPhi, near2, mixedNext, dt, rDt, Sp, Su, sp, ctl);
}

// The same pass as a CAPPED grid ("dense_ctas" CTAs per SM, "dense_threads" threads each) whose CTAs walk the tiles
// interleaved: a fixed share of every SM for the streaming pass for as long as the interface kernels run beside it on
// the other stream, instead of 8 resident CTAs per SM that saturate DRAM (and every warp slot) for 0.43 ms and stretch
// the latency-bound interface kernels.  Same per-cell code, same bits.
__global__ void __launch_bounds__(256, 8) k_dense_update_capped(MeshDev m, const double* __restrict__ aOld, double* __restrict__ aNew,
                                                             const double* __restrict__ phi, const double* __restrict__ alphaB,
                                                             double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2,
                                                             unsigned int* __restrict__ mixedNext, double dt, double rDt,
                                                             const double* __restrict__ Sp, const double* __restrict__ Su,
                                                             StepParams sp, Ctl* ctl, int tile0, int nTiles)
{
    double mn = SV_VGREAT, mx = -SV_VGREAT;
    for (int tile = tile0 + blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const int c = tile * blockDim.x + threadIdx.x;
        const bool mixed = denseCell<0>(m, c, aOld, aNew, phi, alphaB, alphaPhi, near2, dt, rDt, Sp, Su, sp, mn, mx);
        const unsigned int w = __ballot_sync(0xffffffffu, mixed);
        if ((threadIdx.x & 31) == 0 && c < m.nCells) mixedNext[c >> 5] = w;
    }
    blockMinMaxFast(mn, mx, &ctl->minDense, &ctl->maxDense);
}

// K7, owner-sorted layout (the default when the mesh allows it).  Same arithmetic, summation order and outputs as
// k_dense_update; what changes is the connectivity the kernel streams.  In an OpenFOAM mesh the internal faces are
// ordered by owner (upper-triangular order), so
//   * the internal faces a cell OWNS are one contiguous face range: their phi, neighbour and alphaPhi entries are
//     read/written as contiguous runs (no row entries at all: 12 B/cell of `neighbour` instead of 24 B of int2 rows),
//   * the faces it is the NEIGHBOUR of all have lower indices than the ones it owns, so its ascending-face row is
//     [neighbour-side faces ascending] ++ [owned internal ascending] ++ [owned boundary ascending]; only the first part
//     needs row entries {face, owner}: 24 B/cell for a hex instead of 48,
//   * the two row offsets come from one count byte per cell (owned | neighbour-side << 4) and a warp prefix sum on top
//     of one int2 per 32 cells, instead of a 4-byte offset per cell.
// Cells with boundary faces (2.3 % at 256^3) take the generic row path of k_dense_update inside the same kernel.
// Connectivity per hex cell: 12 + 24 + 1 + 0.4 = 37.4 B against 52 B before (algorithmic: 24 B).
struct DenseFast {
    const unsigned char* cnt;      // [nC] owned internal faces | neighbour-side faces << 4
    const int2* warpBase;          // [ceil(nC/32)] {first owned internal face, first neighbour-side row} of the first cell of the warp
    const int2* lowRows;           // [nIF] {face, owner} grouped by neighbour cell, ascending face
    const unsigned int* slowBits;  // [ceil(nC/32)] cells that take the generic row path
    int enabled;
};

__device__ __forceinline__ bool isMixed(double a, double tol) { return (tol < a) && (a < 1.0 - tol); }

#ifndef SV_DCH
#define SV_DCH 3   // faces fetched per batch (independent loads in flight per thread)
#endif
#ifndef SV_DMINB
#define SV_DMINB 8
#endif
__global__ void __launch_bounds__(256, SV_DMINB) k_dense_update2(MeshDev m, DenseFast df, const double* __restrict__ aOld, double* __restrict__ aNew,
                                                         const double* __restrict__ phi, const double* __restrict__ alphaB,
                                                         double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2,
                                                         unsigned int* __restrict__ mixedNext, double dt, double rDt,
                                                         const double* __restrict__ Sp, const double* __restrict__ Su, StepParams sp, Ctl* ctl,
                                                         int nTiles)
{
  // one 256-cell tile per CTA by default; with the "dense_ctas" option a capped grid of CTAs walks the tiles
  // (a fixed share of every SM for the streaming pass while the interface kernels run beside it)
  double mn = SV_VGREAT, mx = -SV_VGREAT;
  const int lane = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int c = tile * blockDim.x + threadIdx.x;
    const int w = c >> 5;
    const bool inRange = c < m.nCells;
    const bool warpIn = (w << 5) < m.nCells;
    const unsigned int cv = inRange ? (unsigned int)__ldg(df.cnt + c) : 0u;
    const unsigned int packed = (cv & 15u) | ((cv >> 4) << 16);
    unsigned int incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned int excl = incl - packed;
    int2 wb = make_int2(0, 0);
    unsigned int slowW = 0u, n2W = 0u;
    if (warpIn) {
        wb = __ldg(df.warpBase + w);
        slowW = __ldg(df.slowBits + w);
        n2W = __ldg(near2 + w);
    }
    bool mixed = false;
    if (inRange && !((n2W >> lane) & 1u)) {
        const double aC = __ldg(aOld + c);
        double sum = 0.0;
        if (!((slowW >> lane) & 1u)) {
            const int nOwn = (int)(cv & 15u), nLow = (int)(cv >> 4);
            const int fOwn = wb.x + (int)(excl & 0xffffu), rLow = wb.y + (int)(excl >> 16);
            // neighbour-side faces: this cell is the neighbour, the owner is upwind when phi >= 0
            for (int j0 = 0; j0 < nLow; j0 += SV_DCH) {
                int2 e[SV_DCH];
                double ph[SV_DCH], aN[SV_DCH];
#pragma unroll
                for (int q = 0; q < SV_DCH; ++q) e[q] = (j0 + q < nLow) ? __ldg(df.lowRows + rLow + j0 + q) : make_int2(0, 0);
#pragma unroll
                for (int q = 0; q < SV_DCH; ++q) {
                    ph[q] = 0.0;
                    aN[q] = 0.0;
                    if (j0 + q < nLow) {
                        ph[q] = __ldg(phi + e[q].x);
                        aN[q] = __ldg(aOld + e[q].y);
                    }
                }
#pragma unroll
                for (int q = 0; q < SV_DCH; ++q) {
                    if (j0 + q < nLow) {
                        const double aUp = (ph[q] >= 0) ? aN[q] : aC;
                        sum -= (ph[q] * aUp) * dt;
                    }
                }
            }
            // owned internal faces: a contiguous face range
            for (int j0 = 0; j0 < nOwn; j0 += SV_DCH) {
                int nb[SV_DCH];
                double ph[SV_DCH], aN[SV_DCH];
#pragma unroll
                for (int q = 0; q < SV_DCH; ++q) {
                    nb[q] = 0;
                    ph[q] = 0.0;
                    if (j0 + q < nOwn) {
                        nb[q] = __ldg(m.neighbour + fOwn + j0 + q);
                        ph[q] = __ldg(phi + fOwn + j0 + q);
                    }
                }
#pragma unroll
                for (int q = 0; q < SV_DCH; ++q) aN[q] = (j0 + q < nOwn) ? __ldg(aOld + nb[q]) : 0.0;
#pragma unroll
                for (int q = 0; q < SV_DCH; ++q) {
                    if (j0 + q < nOwn) {
                        const double aUp = (ph[q] >= 0) ? aC : aN[q];
                        const double dvf = (ph[q] * aUp) * dt;
                        sum += dvf;
                        alphaPhi[fOwn + j0 + q] = divz(dvf, dt);
                    }
                }
            }
        } else {
            const int k0 = __ldg(m.cellOff + c), k1 = __ldg(m.cellOff + c + 1);
            for (int k = k0; k < k1; ++k) {
                const int2 e = __ldg(m.cellAsc + k);
                const int f = e.x & 0x7fffffff;
                const bool flip = e.x < 0;
                const double ph = __ldg(phi + f);
                double aUp;
                if (e.y >= 0) {
                    const bool selfUp = flip ? (ph < 0) : (ph >= 0);
                    aUp = selfUp ? aC : __ldg(aOld + e.y);
                } else {
                    const int bf = -1 - e.y;
                    if (__ldg(m.bKind + bf) == 1) continue;  // empty patch: no field
                    aUp = __ldg(alphaB + bf);
                }
                const double dvf = (ph * aUp) * dt;
                if (!flip) {
                    sum += dvf;
                    alphaPhi[f] = divz(dvf, dt);
                } else {
                    sum -= dvf;
                }
            }
        }
        const double ivf = divz(sum, __ldg(m.V + c));
        double num = aC * rDt;
        if (Su) num = num + Su[c];
        num = num - ivf * rDt;
        double a = divz(num, (Sp ? (rDt - Sp[c]) : rDt));
        mn = dmin(mn, a);
        mx = dmax(mx, a);
        a = snapClip(a, sp.snapTol, sp.clip);
        aNew[c] = a;
        mixed = isMixed(a, sp.mixedTol);
    }
    const unsigned int wbits = __ballot_sync(0xffffffffu, mixed);
    if (lane == 0 && inRange) mixedNext[w] = wbits;
  }
  blockMinMaxFast(mn, mx, &ctl->minDense, &ctl->maxDense);
}

// K7, batched variant (round 2).  Same rows, arithmetic, summation order and outputs as k_dense_update; what changes is the
// shape of the memory traffic.  ncu of k_dense_update: 456 executed instructions per cell, issue slots 50 % busy, one or two
// loads in flight per thread (row entry -> phi -> alpha is a dependent chain walked face by face through a branchy loop).
// Here a cell with six faces and a 16-byte aligned row (every cell of a hexahedral mesh) loads its whole row with three
// 16-byte requests, then issues the six phi gathers back to back, then the (predicated) alpha gathers, and only then does
// the arithmetic in ascending face order: up to six independent loads in flight per thread and no loop-carried branches.
// Other cells take the generic loop.
#ifndef SV_D3MINB
#define SV_D3MINB 5
#endif
__global__ void __launch_bounds__(256, SV_D3MINB) k_dense_update3(MeshDev m, const double* __restrict__ aOld, double* __restrict__ aNew,
                                                                  const double* __restrict__ phi, const double* __restrict__ alphaB,
                                                                  double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2,
                                                                  unsigned int* __restrict__ mixedNext, double dt, double rDt,
                                                                  const double* __restrict__ Sp, const double* __restrict__ Su, StepParams sp,
                                                                  Ctl* ctl)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    double mn = SV_VGREAT, mx = -SV_VGREAT;
    bool mixed = false;
    if (c < m.nCells && !bitTest(near2, c)) {
        const int k0 = __ldg(m.cellOff + c), k1 = __ldg(m.cellOff + c + 1);
        const double aC = __ldg(aOld + c);
        double sum = 0.0;
        if (k1 - k0 == 6 && (k0 & 1) == 0) {
            int ex[6], ey[6];
            {
                const int4* row = reinterpret_cast<const int4*>(m.cellAsc + k0);
                const int4 r0 = __ldg(row), r1 = __ldg(row + 1), r2 = __ldg(row + 2);
                ex[0] = r0.x; ey[0] = r0.y; ex[1] = r0.z; ey[1] = r0.w;
                ex[2] = r1.x; ey[2] = r1.y; ex[3] = r1.z; ey[3] = r1.w;
                ex[4] = r2.x; ey[4] = r2.y; ex[5] = r2.z; ey[5] = r2.w;
            }
            double ph[6], aUp[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) ph[q] = __ldg(phi + (ex[q] & 0x7fffffff));
            bool skip[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const bool flip = ex[q] < 0;
                skip[q] = false;
                if (ey[q] >= 0) {
                    const bool selfUp = flip ? (ph[q] < 0) : (ph[q] >= 0);
                    aUp[q] = aC;
                    if (!selfUp) aUp[q] = __ldg(aOld + ey[q]);
                } else {
                    const int bf = -1 - ey[q];
                    skip[q] = (__ldg(m.bKind + bf) == 1);   // empty patch: no field
                    aUp[q] = __ldg(alphaB + bf);
                }
            }
            double dvf[6];
            bool anyNZ = false;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                dvf[q] = (ph[q] * aUp[q]) * dt;
                anyNZ |= (!skip[q] && dvf[q] != 0.0);
            }
            const bool warpNZ = __any_sync(__activemask(), anyNZ);   // x/dt == x bitwise for x = +-0: skip the divides where the warp is empty
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                if (skip[q]) continue;
                if (ex[q] >= 0) {
                    sum += dvf[q];
                    double ap = dvf[q];
                    if (warpNZ) ap = (ap == 0.0) ? ap : ap / dt;
                    alphaPhi[ex[q]] = ap;
                } else {
                    sum -= dvf[q];
                }
            }
        } else {
            for (int k = k0; k < k1; ++k) {
                const int2 e = __ldg(m.cellAsc + k);
                const int f = e.x & 0x7fffffff;
                const bool flip = e.x < 0;
                const double p = __ldg(phi + f);
                double a;
                if (e.y >= 0) {
                    const bool selfUp = flip ? (p < 0) : (p >= 0);
                    a = selfUp ? aC : __ldg(aOld + e.y);
                } else {
                    const int bf = -1 - e.y;
                    if (__ldg(m.bKind + bf) == 1) continue;
                    a = __ldg(alphaB + bf);
                }
                const double d = (p * a) * dt;
                if (!flip) {
                    sum += d;
                    alphaPhi[f] = divz(d, dt);
                } else {
                    sum -= d;
                }
            }
        }
        const double ivf = divz(sum, __ldg(m.V + c));
        double num = aC * rDt;
        if (Su) num = num + Su[c];
        num = num - ivf * rDt;
        double a = divz(num, (Sp ? (rDt - Sp[c]) : rDt));
        mn = a;
        mx = a;
        a = snapClip(a, sp.snapTol, sp.clip);
        aNew[c] = a;
        mixed = (sp.mixedTol < a) && (a < 1.0 - sp.mixedTol);
    }
    const unsigned int w = __ballot_sync(0xffffffffu, mixed);
    if ((threadIdx.x & 31) == 0 && c < m.nCells) mixedNext[c >> 5] = w;
    blockMinMaxFast(mn, mx, &ctl->minDense, &ctl->maxDense);
}

// K7 over a sliced, transposed copy of the rows (round 2).  k_dense_update reads its int2 rows with a 48-byte stride between
// the lanes of a warp: every load instruction touches 12 lines for 256 useful bytes and relies on L1 to keep them for the
// next five iterations (ncu: 63 % of the L2->L1 sectors "excessive", L1 hit rate 52 %).  Here the rows of each group of 32
// consecutive cells are stored entry-major -- entry q of the 32 cells contiguous -- so each row load is one fully used
// 256-byte request and L1 is left to the phi / alpha gathers.  Same arithmetic, summation order and outputs; the loop trip
// count is the widest row of the slice (warp uniform), shorter rows are padded with a sentinel.
#define SV_ROW_PAD ((int)0x80000000)   // {PAD, PAD}: e.y == INT_MIN is no cell and no boundary face
struct DenseSliced {
    const int2* rowsT;     // [sliceOff[nSlices] * 32]
    const int* sliceOff;   // [nSlices + 1], in units of 32 entries
    int enabled;
};
__global__ void __launch_bounds__(256, 8) k_dense_update4(MeshDev m, DenseSliced ds, const double* __restrict__ aOld, double* __restrict__ aNew,
                                                          const double* __restrict__ phi, const double* __restrict__ alphaB,
                                                          double* __restrict__ alphaPhi, const unsigned int* __restrict__ near2,
                                                          unsigned int* __restrict__ mixedNext, double dt, double rDt,
                                                          const double* __restrict__ Sp, const double* __restrict__ Su, StepParams sp,
                                                          Ctl* ctl)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double mn = SV_VGREAT, mx = -SV_VGREAT;
    bool mixed = false;
    const bool inRange = c < m.nCells;
    const int slice = c >> 5;
    int b0 = 0, b1 = 0;
    if (((slice << 5) < m.nCells)) {
        b0 = __ldg(ds.sliceOff + slice);
        b1 = __ldg(ds.sliceOff + slice + 1);
    }
    const bool work = inRange && !bitTest(near2, c);
    const double aC = inRange ? __ldg(aOld + c) : 0.0;
    double sum = 0.0;
    for (int q = b0; q < b1; ++q) {
        const int2 e = __ldg(ds.rowsT + ((size_t)q << 5) + lane);
        if (!work || (e.x == SV_ROW_PAD && e.y == SV_ROW_PAD)) continue;
        const int f = e.x & 0x7fffffff;
        const bool flip = e.x < 0;
        const double ph = __ldg(phi + f);
        double aUp;
        if (e.y >= 0) {
            const bool selfUp = flip ? (ph < 0) : (ph >= 0);
            aUp = selfUp ? aC : __ldg(aOld + e.y);
        } else {
            const int bf = -1 - e.y;
            if (__ldg(m.bKind + bf) == 1) continue;  // empty patch: no field
            aUp = __ldg(alphaB + bf);
        }
        const double dvf = (ph * aUp) * dt;
        if (!flip) {
            sum += dvf;
            alphaPhi[f] = divz(dvf, dt);
        } else {
            sum -= dvf;
        }
    }
    if (work) {
        const double ivf = divz(sum, __ldg(m.V + c));
        double num = aC * rDt;
        if (Su) num = num + Su[c];
        num = num - ivf * rDt;
        double a = divz(num, (Sp ? (rDt - Sp[c]) : rDt));
        mn = a;
        mx = a;
        a = snapClip(a, sp.snapTol, sp.clip);
        aNew[c] = a;
        mixed = (sp.mixedTol < a) && (a < 1.0 - sp.mixedTol);
    }
    const unsigned int w = __ballot_sync(0xffffffffu, mixed);
    if (lane == 0 && inRange) mixedNext[c >> 5] = w;
    blockMinMaxFast(mn, mx, &ctl->minDense, &ctl->maxDense);
}

// K7, staged variant: the same arithmetic as k_dense_update, but the CTA first copies (cp.async, 16-byte
// requests, no registers) its contiguous slab of cell->face rows and the phi values of the faces its cells
// OWN (contiguous too: OpenFOAM orders internal faces by owner) into shared memory.  That turns the two
// biggest gathers (48 B/cell of rows read with a 48-byte stride, 12 B/cell... of owned phi with a 24-byte
// stride) into fully coalesced bulk copies with ~20 KB in flight per CTA, instead of one dependent
// chain of three loads per face per thread.
struct DenseStage {
    const int* ctaFace;   // [nCta+1] first owned INTERNAL face of the first cell of each CTA (nullptr: no phi staging)
    int rowCap;           // int2 entries of row storage in shared memory
    int phiCap;           // doubles of phi storage
};

__device__ __forceinline__ void cpAsync16(void* smem, const void* gmem)
{
    const unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__global__ void __launch_bounds__(256, 6) k_dense_update_staged(MeshDev m, DenseStage ds, const double* __restrict__ aOld,
                                                                double* __restrict__ aNew, const double* __restrict__ phi,
                                                                const double* __restrict__ alphaB, double* __restrict__ alphaPhi,
                                                                const unsigned int* __restrict__ near2, unsigned int* __restrict__ mixedNext,
                                                                double dt, double rDt, const double* __restrict__ Sp,
                                                                const double* __restrict__ Su, StepParams sp, Ctl* ctl)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    int2* rows = reinterpret_cast<int2*>(dsm);                                   // [rowCap]  (+2 for alignment slack)
    double* sphi = reinterpret_cast<double*>(dsm + (size_t)(ds.rowCap + 2) * 8);  // [phiCap]  (+2)
    const int c0 = blockIdx.x * blockDim.x;
    const int c = c0 + threadIdx.x;
    const int cEnd = min(c0 + (int)blockDim.x, m.nCells);
    // slab of rows [rowBeg, rowEnd) and owned internal faces [fBeg, fEnd)
    const int rowBeg = __ldg(m.cellOff + c0), rowEnd = __ldg(m.cellOff + cEnd);
    const int rowBegA = rowBeg & ~1;  // 16-byte aligned start (int2 = 8 B)
    const int nRow16 = (rowEnd - rowBegA + 1) >> 1;
    const bool rowsStaged = (rowEnd - rowBegA) <= ds.rowCap;
    if (rowsStaged) {
        for (int i = threadIdx.x; i < nRow16; i += blockDim.x) cpAsync16(rows + 2 * i, m.cellAsc + rowBegA + 2 * i);
    }
    int fBeg = 0, fEnd = 0, fBegA = 0;
    bool phiStaged = false;
    if (ds.ctaFace) {
        fBeg = __ldg(ds.ctaFace + blockIdx.x);
        fEnd = __ldg(ds.ctaFace + blockIdx.x + 1);
        fBegA = fBeg & ~1;
        phiStaged = (fEnd - fBegA) <= ds.phiCap;
        if (phiStaged) {
            const int n16 = (fEnd - fBegA + 1) >> 1;
            // the tail request may read one double past fEnd: phi has nF >= nIF + ... entries and is allocated with slack
            for (int i = threadIdx.x; i < n16; i += blockDim.x) cpAsync16(sphi + 2 * i, phi + fBegA + 2 * i);
        }
    }
    // independent loads while the copies are in flight
    const bool inRange = c < m.nCells;
    const bool work = inRange && !bitTest(near2, c);
    const int k0 = inRange ? __ldg(m.cellOff + c) : 0, k1 = inRange ? __ldg(m.cellOff + c + 1) : 0;
    const double aC = inRange ? __ldg(aOld + c) : 0.0;
    const double Vc = inRange ? __ldg(m.V + c) : 1.0;
    cpAsyncWaitAll();
    __syncthreads();

    double mn = SV_VGREAT, mx = -SV_VGREAT;
    bool mixed = false;
    if (work) {
        double sum = 0.0;
        for (int k = k0; k < k1; ++k) {
            const int2 e = rowsStaged ? rows[k - rowBegA] : __ldg(m.cellAsc + k);
            const int f = e.x & 0x7fffffff;
            const bool flip = e.x < 0;
            const double ph = (phiStaged && f >= fBeg && f < fEnd) ? sphi[f - fBegA] : __ldg(phi + f);
            double aUp;
            if (e.y >= 0) {
                const bool selfUp = flip ? (ph < 0) : (ph >= 0);
                aUp = selfUp ? aC : __ldg(aOld + e.y);
            } else {
                const int bf = -1 - e.y;
                const unsigned char kind = __ldg(m.bKind + bf);
                if (kind == 1) continue;  // empty patch: no field
                aUp = __ldg(alphaB + bf);
            }
            const double dvf = (ph * aUp) * dt;
            if (!flip) {
                sum += dvf;
                alphaPhi[f] = divz(dvf, dt);
            } else {
                sum -= dvf;
            }
        }
        const double ivf = divz(sum, Vc);
        double num = aC * rDt;
        if (Su) num = num + Su[c];
        num = num - ivf * rDt;
        double a = divz(num, (Sp ? (rDt - Sp[c]) : rDt));
        mn = a;
        mx = a;
        a = snapClip(a, sp.snapTol, sp.clip);
        aNew[c] = a;
        mixed = (sp.mixedTol < a) && (a < 1.0 - sp.mixedTol);
    }
    const unsigned int w = __ballot_sync(0xffffffffu, mixed);
    if ((threadIdx.x & 31) == 0 && inRange) mixedNext[c >> 5] = w;
    blockMinMaxFast(mn, mx, &ctl->minDense, &ctl->maxDense);
}

// true transport of face f seen from cell c: the geometric value where the face is downwind of a
// cut cell (advection.C:134-166,185-217), else the upwind value
__device__ __forceinline__ bool faceDVf(const MeshDev& m, int c, int f, bool flip, int other, double ph,
                                        const double* __restrict__ aOld, const double* __restrict__ alphaB, double dt,
                                        const int* __restrict__ cellSlot, const int* __restrict__ cellStatus,
                                        const double* __restrict__ dVfGeo, double& dvf)
{
    int up;
    if (other >= 0) {
        const int own = flip ? other : c, nei = flip ? c : other;
        up = (ph >= 0) ? own : nei;
    } else {
        if (m.bKind[-1 - other] == 1) return false;
        up = (ph >= 0) ? c : -1;
    }
    if (up >= 0) {
        const int slot = cellSlot[up];
        if (slot >= 0 && cellStatus[slot] == 0) {
            dvf = dVfGeo[f];
            return true;
        }
    }
    return upwindDVf(m, c, f, flip, other, ph, aOld, alphaB, dt, dvf);
}

// the two out-of-bounds tests of the reference are NOT the same expression:
//   limitFlux loop condition (advectionTemplates.C:146): max(alpha) - 1 > aTol || min(alpha) < -aTol
//   boundFlux eligibility     (advectionTemplates.C:245): alpha < -aTol || alpha > 1 + aTol
__device__ __forceinline__ bool oobGlobal(double a) { return ((a - 1.0) > SV_ATOL) || (a < -SV_ATOL); }
__device__ __forceinline__ bool oobBound(double a) { return (a < -SV_ATOL) || (a > 1.0 + SV_ATOL); }

// A10 for the near2 cells (same expression and order as k_dense_update) + dVf scratch for bounding
// + the out-of-bounds list of the first bounding sweep.
__global__ void __launch_bounds__(128) k_near_update(MeshDev m, const int* near2List, const unsigned int* near1, Ctl* ctl,
                                                     const double* __restrict__ aOld, double* aNew, const double* __restrict__ phi,
                                                     const double* __restrict__ alphaB, const int* cellSlot, const int* cellStatus,
                                                     const double* dVfGeo, double* dVf, double dt, double rDt, const double* Sp,
                                                     const double* Su, int* oobList0, unsigned char* oobState)
{
    const int n = ctl->nNear2;
    double mn = SV_VGREAT, mx = -SV_VGREAT;
    int nOobG = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = near2List[i];
        double sum = 0.0;
        for (int k = m.cellOff[c]; k < m.cellOff[c + 1]; ++k) {
            const int2 e = m.cellAsc[k];
            const int f = e.x & 0x7fffffff;
            const bool flip = e.x < 0;
            double dvf;
            if (!faceDVf(m, c, f, flip, e.y, phi[f], aOld, alphaB, dt, cellSlot, cellStatus, dVfGeo, dvf)) continue;
            dVf[f] = dvf;  // both sides store the identical value
            if (!flip) sum += dvf; else sum -= dvf;
        }
        const double ivf = sum / m.V[c];
        double num = aOld[c] * rDt;
        if (Su) num = num + Su[c];
        num = num - ivf * rDt;
        const double a = num / (Sp ? (rDt - Sp[c]) : rDt);
        aNew[c] = a;
        mn = dmin(mn, a);
        mx = dmax(mx, a);
        if (oobGlobal(a)) nOobG++;
        if (oobBound(a) && bitTest(near1, c)) {  // needBounding = near1 (advectionTemplates.C:138-141)
            oobState[c] = 1;
            oobList0[atomicAdd(&ctl->nOob[0], 1)] = c;
        }
    }
    if (nOobG) atomicAdd(&ctl->nearOob[0], nOobG);
    blockMinMax(mn, mx, &ctl->minNear0, &ctl->maxNear0);
}

// ---- A11: limitFlux / boundFlux (advectionTemplates.C:118-349) -----------------------------------
// Work is proportional to the number of out-of-bounds cells, not to the mesh:
//   sweep 0 starts from the list k_near_update built; sweep s+1 from the list k_bound_apply(s) built
//   (a cell can only go/stay out of bounds where a correction touched it).
// The sweeps do not depend on the streaming kernel: if the global loop condition holds only because
// of cells outside needBounding the reference's sweep is a no-op too, so the sweeps run
// unconditionally and "number of sweeps executed" is derived afterwards from the recorded counts.
// Scratch arrays are validated by a per-sweep tag instead of being cleared.
#ifdef SV_BOUND_STATS
__device__ unsigned long long g_dbg[8];
#endif
// per-(step, sweep) validity tag of the bounding scratch arrays
__device__ __forceinline__ int boundTag(const Ctl* ctl, int s) { return ctl->epoch * (SV_MAX_SWEEPS + 1) + s + 1; }
struct BoundScratch {
    double* corr;   // dVfCorrectionValues, valid where tagV == tag
    int* tagV;
    int* tagR;      // == tag where the face was recorded in correctedFaces (first inner iteration)
    int* corrBy;    // ... by which cell
    int* corrPos;   // ... at which position of that cell's list
    int* affStamp;  // per cell: == tag once the cell is in the affected list
};

__device__ __forceinline__ bool faceActive(const MeshDev& m, int f) { return f < m.nIF || m.bKind[f - m.nIF] != 1; }
__device__ __forceinline__ double corrVal(const BoundScratch& b, int f, int tag) { return (b.tagV[f] == tag) ? b.corr[f] : 0.0; }

// ---- one cell of boundFlux (advectionTemplates.C:245-346) ---------------------------------------
// These kernels are pure latency: a few thousand threads, each a chain of dependent gathers.  The
// cell's faces are therefore fetched in unrolled batches of 8 (row -> 8 faces -> 8x6 independent
// gathers -> local arrays), i.e. ~4 memory round trips per cell instead of ~10 per face per inner
// iteration, and the inner iterations then run on thread-local data only.  Corrections are written
// back at the end: a face is only ever corrected by its upwind cell, so the local copy is exact.
// MAXBF (faces per cell the bounding kernels keep in thread-local arrays) is a template parameter chosen from
// the mesh: with 64 the 3.6 KB/thread of local memory thrashed L1 (measured 25k cycles per cell for ONE
// inner iteration); a hex mesh uses 8.
// CellBound doubles as the per-out-of-bounds-cell RECORD that k_bound_deps (parallel over all cells)
// writes to global memory, so that the latency-critical chain walk of k_bound_run costs one contiguous
// record read per cell instead of three dependent gathers.
template <int SV_MAXBF>
struct __align__(16) CellBound {
    double fPhi[SV_MAXBF], fDvf[SV_MAXBF], fCorr[SV_MAXBF];
    double V, alpha, aOld, Sp, Su;
    unsigned long long ownMask, downMask;
    unsigned long long predMask;  // faces whose other cell is a predecessor (lower index, out of bounds, upwind of this cell)
    unsigned long long succMask;  // faces whose other cell is a successor   (higher index, out of bounds, this cell upwind)
    int fId[SV_MAXBF], other[SV_MAXBF];
    int nf, pad_;
};

template <int SV_MAXBF>
// phiBits / phiHost / phiDev (svof_step_host, zero-copy form): the device holds phi only on the faces its bitmap marks (a
// cell with alpha != 0 on one side).  An out-of-bounds cell whose own alpha WAS zero -- it received a negative sliver from a
// neighbour, or was overfilled at Courant > 1 -- has unmarked faces: their entries are read here straight from the caller's
// pinned phi, stored and marked, so the sweep (and the alphaPhi read-back) sees exactly the full field.
__device__ __forceinline__ void loadCellBound(const MeshDev& m, int celli, const double* phi, const double* dVf,
                                              const BoundScratch& b, int tag, CellBound<SV_MAXBF>& cb, unsigned int* phiBits = nullptr,
                                              const double* phiHost = nullptr, double* phiDev = nullptr, Ctl* ctl = nullptr)
{
    const int c0 = m.cellOff[celli];
    int nf = m.cellOff[celli + 1] - c0;
    if (nf > SV_MAXBF) nf = SV_MAXBF;
    cb.nf = nf;
    cb.ownMask = cb.downMask = cb.predMask = cb.succMask = 0ull;
    for (int q0 = 0; q0 < nf; q0 += 8) {
        int ff[8], ow[8], nb[8], tg[8];
        double ph[8], dv[8], cr[8];
        unsigned char bk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ff[j] = (q0 + j < nf) ? m.cellFaces[c0 + q0 + j] : -1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ow[j] = 0; nb[j] = -1; tg[j] = 0; ph[j] = 0.0; dv[j] = 0.0; cr[j] = 0.0; bk[j] = 0;
            if (ff[j] >= 0) {
                const int f = ff[j];
                ow[j] = m.owner[f];
                if (phiHost && !((__ldcg(phiBits + (f >> 5)) >> (f & 31)) & 1u)) {
                    ph[j] = *(const volatile double*)(phiHost + f);
                    phiDev[f] = ph[j];
                    __threadfence();   // whoever sees the bit sees the value
                    atomicOr(phiBits + (f >> 5), 1u << (f & 31));
                    atomicAdd(&ctl->nPhiPulled, 1);
                } else {
                    ph[j] = __ldcg(phi + f);
                }
                dv[j] = dVf[f];
                tg[j] = __ldcg(b.tagV + f);  // L2 reads: written by other SMs during this kernel
                cr[j] = __ldcg(b.corr + f);
                if (f < m.nIF) nb[j] = m.neighbour[f]; else bk[j] = m.bKind[f - m.nIF];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (ff[j] < 0) continue;
            const int q = q0 + j;
            const bool act = (bk[j] != 1);                     // faceValue(): empty patches read as zero
            const bool isOwn = (ow[j] == celli);
            const double p = act ? ph[j] : 0.0;
            const bool down = isOwn ? (p >= 0) : (p < 0);      // setDownwindFaces, advection.C:242-252
            const int oth = isOwn ? nb[j] : ow[j];
            double c = (act && tg[j] == tag) ? cr[j] : 0.0;
            // a correction on a face that is downwind of the OTHER cell was written by that cell; in the
            // ascending sweep of the reference this cell has only seen it if the writer has a lower index
            if (act && !down && oth > celli) c = 0.0;
            cb.fId[q] = ff[j];
            cb.other[q] = act ? oth : -1;
            cb.fPhi[q] = p;
            cb.fDvf[q] = act ? dv[j] : 0.0;
            cb.fCorr[q] = c;
            if (isOwn) cb.ownMask |= 1ull << q;
            if (down) cb.downMask |= 1ull << q;
        }
    }
}

template <int SV_MAXBF>
__device__ bool boundCell(int celli, CellBound<SV_MAXBF>& cb, const BoundScratch& b, int tag, double dt, double rDeltaT)
{
    // Small records (hex, prism, tet: MAXBF <= 16) are walked with FULLY UNROLLED, predicated face loops so that the
    // record and the scratch arrays live in registers: k_bound_run is a chain of single-thread evaluations, and with
    // dynamically indexed (local-memory) arrays one cell cost 9.8k cycles (measured), most of it exposed load latency.
    constexpr int UNR = SV_MAXBF <= 16 ? SV_MAXBF : 1;
    bool hadRoom = false;
    const double Vi = cb.V;
    const int nf = cb.nf;
    const int qEnd = SV_MAXBF <= 16 ? SV_MAXBF : nf;
    double room[SV_MAXBF];
    int recPos[SV_MAXBF];
    unsigned long long modMask = 0, recMask = 0;
    const double a0 = cb.alpha;
    const double SuI = cb.Su, SpI = cb.Sp;
    const double aOldI = cb.aOld;
    double alphaOvershoot = pos0(a0 - 1.0) * (a0 - 1.0) + neg0(a0) * a0;
    double fluidToPassOn = alphaOvershoot * Vi;
    int nFacesToPassFluidThrough = 1;
    bool firstLoop = true;
    int nRecorded = 0;
    for (int iter = 0; iter < 10; ++iter) {
        if (fabs(alphaOvershoot) < SV_ATOL || nFacesToPassFluidThrough == 0) break;
#ifdef SV_BOUND_STATS
        atomicAdd(&g_dbg[1], 1ull);
#endif
        // facesToPassFluidThrough / dVfmax / dVftot: fixed before any correction of this iteration
        double dVftot = 0;
        nFacesToPassFluidThrough = 0;
#pragma unroll UNR
        for (int q = 0; q < qEnd; ++q) {
            if (q >= nf) continue;
            double r = -1.0;
            if ((cb.downMask >> q) & 1ull) {
                const double dVff = cb.fDvf[q] + cb.fCorr[q];
                const double maxExtra = fabs(pos0(fluidToPassOn) * cb.fPhi[q] * dt - dVff);
                if (maxExtra / Vi > SV_ATOL) {
                    r = maxExtra;
                    dVftot += fabs(cb.fPhi[q] * dt);
                }
            }
            room[q] = r;
        }
#pragma unroll UNR
        for (int q = 0; q < qEnd; ++q) {
            if (q >= nf || room[q] < 0.0) continue;
            double through = fabs(fluidToPassOn) * fabs(cb.fPhi[q] * dt) / dVftot;
            nFacesToPassFluidThrough += int(pos0(room[q] - through));
            through = dmin(through, room[q]);
            double dVff = cb.fCorr[q];
            dVff += sgn(cb.fPhi[q]) * sgn(fluidToPassOn) * through;
            cb.fCorr[q] = dVff;
            modMask |= 1ull << q;
            hadRoom = true;
            if (firstLoop) {
                recMask |= 1ull << q;
                recPos[q] = nRecorded++;
            }
        }
        firstLoop = false;
        double nfl = 0.0, nc = 0.0;  // netFlux(dVf_), netFlux(dVfCorrectionValues)  (advection.C:259-288)
#pragma unroll UNR
        for (int q = 0; q < qEnd; ++q) {
            if (q >= nf) continue;
            if ((cb.ownMask >> q) & 1ull) {
                nfl += cb.fDvf[q];
                nc += cb.fCorr[q];
            } else {
                nfl -= cb.fDvf[q];
                nc -= cb.fCorr[q];
            }
        }
        const double alpha1New = (aOldI * rDeltaT + SuI - nfl / Vi * rDeltaT - nc / Vi * rDeltaT) / (rDeltaT - SpI);
        alphaOvershoot = pos0(alpha1New - 1.0) * (alpha1New - 1.0) + neg0(alpha1New) * alpha1New;
        fluidToPassOn = alphaOvershoot * Vi;
    }
#pragma unroll UNR
    for (int q = 0; q < qEnd; ++q) {
        if (q >= nf || !((modMask >> q) & 1ull)) continue;
        const int f = cb.fId[q];
        b.corr[f] = cb.fCorr[q];
        b.tagV[f] = tag;
        if ((recMask >> q) & 1ull) {
            b.corrBy[f] = celli;
            b.corrPos[f] = recPos[q];
            b.tagR[f] = tag;
        }
    }
    return hadRoom;
}

// The reference sweeps the cells in ascending index (Gauss-Seidel, SURVEY 8a' item 15).  A cell only
// ever READS corrections written by another cell on the faces that are DOWNWIND of that other cell,
// so cell y depends exactly on the lower-index out-of-bounds neighbours that are upwind of it across
// the shared face; everything else commutes (reads of higher-index writers are masked in
// loadCellBound).  The sweep is therefore a DAG executed by dependency counting:
//   k_bound_deps : per out-of-bounds cell, count its predecessors (and mark the affected set)
//   k_bound_run  : roots start at once; whoever finishes the last predecessor of a cell runs that
//                  cell next (continuation passing) -- no polling, no single-CTA drain loop.
template <int SV_MAXBF>
__global__ void __launch_bounds__(128) k_bound_deps(MeshDev m, Ctl* ctl, int s, const int* oobList,
                                                    const unsigned char* oobState, const double* alpha, const double* aOld,
                                                    double* phi, const double* dVf, const double* Sp,
                                                    const double* Su, BoundScratch b, int* depInit, int* depLeft, int* oobIdx,
                                                    CellBound<SV_MAXBF>* recs, int capRec, int* affList,
                                                    unsigned int* phiBits = nullptr, const double* phiHost = nullptr)
{
    const int n = ctl->nOob[s & 1];
    const int tag = boundTag(ctl, s);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = oobList[i];
        CellBound<SV_MAXBF> cb;
        loadCellBound(m, c, phi, dVf, b, tag, cb, phiBits, phiHost, phi, ctl);  // corrections of this sweep do not exist yet: fCorr == 0
        if (phiBits && !phiHost) {   // staged svof_step_host: phi was uploaded only where a neighbouring cell held liquid: every face of this cell must be among them
            for (int q = 0; q < cb.nf; ++q)
                if (!((phiBits[cb.fId[q] >> 5] >> (cb.fId[q] & 31)) & 1u)) ctl->phiUnsafe = 1;
        }
        cb.V = m.V[c];
        cb.alpha = alpha[c];
        cb.aOld = aOld[c];
        cb.Sp = Sp ? Sp[c] : 0.0;
        cb.Su = Su ? Su[c] : 0.0;
        int newIds[SV_MAXBF + 1];
        int nNew = 0, deps = 0;
        if (atomicExch(&b.affStamp[c], tag) != tag) newIds[nNew++] = c;
        for (int q0 = 0; q0 < cb.nf; q0 += 8) {
            unsigned char st[8];
            int old[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                st[j] = 0; old[j] = tag;
                const int q = q0 + j;
                if (q < cb.nf && cb.other[q] >= 0) {
                    st[j] = oobState[cb.other[q]];
                    old[j] = atomicExch(&b.affStamp[cb.other[q]], tag);  // affected set = cell + face neighbours
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = q0 + j;
                if (q >= cb.nf || cb.other[q] < 0) continue;
                const int y = cb.other[q];
                if (old[j] != tag) newIds[nNew++] = y;
                if (st[j] != 1) continue;
                const bool down = (cb.downMask >> q) & 1ull;
                if (y < c && !down) { cb.predMask |= 1ull << q; deps++; }
                if (y > c && down) cb.succMask |= 1ull << q;
            }
        }
        depInit[c] = deps;
        depLeft[c] = deps;
        oobIdx[c] = i;
        if (i < capRec) recs[i] = cb; else atomicOr(&ctl->err, SVERR_LIST);
        if (nNew) {
            const int pos = atomicAdd(&ctl->nAff[s], nNew);
            for (int j = 0; j < nNew; ++j) affList[pos + j] = newIds[j];
        }
    }
}

#define SV_BSTACK 64
template <int SV_MAXBF>
__global__ void __launch_bounds__(64) k_bound_run(Ctl* ctl, int s, const int* oobList, unsigned char* oobState, BoundScratch b,
                                                  const int* depInit,
                                                  int* depLeft, const int* oobIdx, const CellBound<SV_MAXBF>* recs, int capRec,
                                                  double dt, double rDt)
{
    const int n = min(ctl->nOob[s & 1], capRec);
    const int tag = boundTag(ctl, s);
    // ONE chain walker per WARP (lane 0): with a walker in every lane the 32 chains of a warp sit at different stages of
    // the load / bound / release cycle and the SIMT serialisation of those divergent sections cost 8.3k cycles of "compute"
    // per cell for one inner iteration (SV_BOUND_STATS, round 2: 4.6k with this form); the work is a few thousand cells,
    // so warps are not scarce
    if (threadIdx.x & 31) return;
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i0 < n; i0 += (gridDim.x * blockDim.x) >> 5) {
        int c = oobList[i0];
        if (depInit[c] != 0) continue;  // released later by whoever finishes its last predecessor
        int i = i0;
        int stack[SV_BSTACK];
        int sp = 0;
        for (;;) {
#ifdef SV_BOUND_STATS
            const long long t0 = clock64();
#endif
            CellBound<SV_MAXBF> cb = recs[i];
            // corrections its predecessors wrote (they are complete: this cell was released by the last of them)
            if (SV_MAXBF <= 16) {
#pragma unroll
                for (int q = 0; q < (SV_MAXBF <= 16 ? SV_MAXBF : 1); ++q) {
                    if (!((cb.predMask >> q) & 1ull)) continue;
                    const int f = cb.fId[q];
                    cb.fCorr[q] = (__ldcg(b.tagV + f) == tag) ? __ldcg(b.corr + f) : 0.0;
                }
            } else {
                for (unsigned long long pm = cb.predMask; pm; pm &= pm - 1) {
                    const int q = __ffsll((long long)pm) - 1;
                    const int f = cb.fId[q];
                    cb.fCorr[q] = (__ldcg(b.tagV + f) == tag) ? __ldcg(b.corr + f) : 0.0;
                }
            }
#ifdef SV_BOUND_STATS
            const long long t1 = clock64();
#endif
            // A cell whose first inner iteration finds no downwind face with room writes nothing, and will write
            // nothing in later sweeps either until a neighbour's correction changes its alpha (its downwind faces are
            // only ever corrected by itself): mark it dormant so later sweeps skip it (exactly the reference's result).
            if (!boundCell(c, cb, b, tag, dt, rDt)) oobState[c] = 3;
#ifdef SV_BOUND_STATS
            const long long t2 = clock64();
            atomicAdd(&g_dbg[0], 1ull);
            atomicAdd(&g_dbg[2], (unsigned long long)(t1 - t0));
            atomicAdd(&g_dbg[3], (unsigned long long)(t2 - t1));
#endif
            if (cb.succMask) {
                __threadfence();  // corrections visible before any successor is released
                if (SV_MAXBF <= 16) {
#pragma unroll
                    for (int q = 0; q < (SV_MAXBF <= 16 ? SV_MAXBF : 1); ++q) {
                        if (!((cb.succMask >> q) & 1ull)) continue;
                        const int y = cb.other[q];
                        if (atomicSub(&depLeft[y], 1) == 1) {  // last predecessor: run y next on this thread
                            if (sp < SV_BSTACK) stack[sp++] = y; else atomicOr(&ctl->err, SVERR_LIST);
                        }
                    }
                } else {
                    for (unsigned long long sm = cb.succMask; sm; sm &= sm - 1) {
                        const int q = __ffsll((long long)sm) - 1;
                        const int y = cb.other[q];
                        if (atomicSub(&depLeft[y], 1) == 1) {
                            if (sp < SV_BSTACK) stack[sp++] = y; else atomicOr(&ctl->err, SVERR_LIST);
                        }
                    }
                }
            }
#ifdef SV_BOUND_STATS
            atomicAdd(&g_dbg[5], (unsigned long long)(clock64() - t2));
#endif
            if (sp == 0) break;
            c = stack[--sp];
            i = oobIdx[c];
            __threadfence();  // acquire side of the release above
        }
    }
}

// k_bound_run for cells with at most 8 faces (hexahedra, prisms, tetrahedra): the same dependency-counted walk, but the
// eight faces of the cell being bounded sit in eight LANES of the walker's warp.  boundFlux's per-face work (two FP64
// divides per face and inner iteration, advectionTemplates.C:285-300) runs in parallel; its ordered sums (dVftot, the two
// netFlux sums of advection.C:259-288) are re-done by every lane from shuffles in face order, so the bits are those of
// the sequential form.  Measured per cell (SV_BOUND_STATS): 8.3k cycles with a walker per lane, 4.6k with one walker per
// warp doing the eight faces in sequence.
__global__ void __launch_bounds__(64) k_bound_run8(Ctl* ctl, int s, const int* oobList, unsigned char* oobState, BoundScratch b,
                                                   const int* depInit, int* depLeft, const int* oobIdx, const CellBound<8>* recs, int capRec,
                                                   double dt, double rDt)
{
    const int n = min(ctl->nOob[s & 1], capRec);
    const int tag = boundTag(ctl, s);
    const int q = threadIdx.x & 31;
    if (q >= 8) return;
    const unsigned M = 0xFFu;
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i0 < n; i0 += (gridDim.x * blockDim.x) >> 5) {
        int c = oobList[i0];
        if (depInit[c] != 0) continue;  // released later by whoever finishes its last predecessor
        int i = i0;
        int stack[SV_BSTACK];           // identical in the eight lanes
        int sp = 0;
        for (;;) {
            const CellBound<8>* rec = recs + i;
            const int nf = rec->nf;
            const bool valid = q < nf;
            const double fPhi = rec->fPhi[q], fDvf = rec->fDvf[q];
            double fCorr = rec->fCorr[q];
            const int fId = rec->fId[q], other = rec->other[q];
            const double Vi = rec->V, a0 = rec->alpha, aOldI = rec->aOld, SpI = rec->Sp, SuI = rec->Su;
            const unsigned long long ownMask = rec->ownMask, downMask = rec->downMask, predMask = rec->predMask, succMask = rec->succMask;
            const bool own = (ownMask >> q) & 1ull, down = (downMask >> q) & 1ull;
            // corrections its predecessors wrote (they are complete: this cell was released by the last of them)
            if (valid && ((predMask >> q) & 1ull)) fCorr = (__ldcg(b.tagV + fId) == tag) ? __ldcg(b.corr + fId) : 0.0;
            // ---- boundFlux for this cell (advectionTemplates.C:245-346)
            bool hadRoom = false, modMine = false, recMine = false;
            int recPosMine = 0;
            double alphaOvershoot = pos0(a0 - 1.0) * (a0 - 1.0) + neg0(a0) * a0;
            double fluidToPassOn = alphaOvershoot * Vi;
            int nFacesToPassFluidThrough = 1;
            bool firstLoop = true;
            for (int iter = 0; iter < 10; ++iter) {
                if (fabs(alphaOvershoot) < SV_ATOL || nFacesToPassFluidThrough == 0) break;
                double room = -1.0, contrib = 0.0;
                if (valid && down) {
                    const double dVff = fDvf + fCorr;
                    const double maxExtra = fabs(pos0(fluidToPassOn) * fPhi * dt - dVff);
                    if (maxExtra / Vi > SV_ATOL) {
                        room = maxExtra;
                        contrib = fabs(fPhi * dt);
                    }
                }
                const unsigned roomMask = __ballot_sync(M, room >= 0.0) & M;
                double dVftot = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double v = __shfl_sync(M, contrib, k);
                    if ((roomMask >> k) & 1u) dVftot += v;
                }
                bool fits = false;
                if (room >= 0.0) {
                    double through = fabs(fluidToPassOn) * fabs(fPhi * dt) / dVftot;
                    fits = pos0(room - through) != 0.0;
                    through = dmin(through, room);
                    double dVff = fCorr;
                    dVff += sgn(fPhi) * sgn(fluidToPassOn) * through;
                    fCorr = dVff;
                    modMine = true;
                    if (firstLoop) {
                        recMine = true;
                        recPosMine = __popc(roomMask & ((1u << q) - 1u));
                    }
                }
                nFacesToPassFluidThrough = __popc(__ballot_sync(M, fits) & M);
                if (roomMask) hadRoom = true;
                firstLoop = false;
                double nfl = 0.0, nc = 0.0;  // netFlux(dVf_), netFlux(dVfCorrectionValues)  (advection.C:259-288)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double vd = __shfl_sync(M, fDvf, k), vc = __shfl_sync(M, fCorr, k);
                    if (k < nf) {
                        if ((ownMask >> k) & 1ull) {
                            nfl += vd;
                            nc += vc;
                        } else {
                            nfl -= vd;
                            nc -= vc;
                        }
                    }
                }
                const double alpha1New = (aOldI * rDt + SuI - nfl / Vi * rDt - nc / Vi * rDt) / (rDt - SpI);
                alphaOvershoot = pos0(alpha1New - 1.0) * (alpha1New - 1.0) + neg0(alpha1New) * alpha1New;
                fluidToPassOn = alphaOvershoot * Vi;
            }
            if (modMine) {
                b.corr[fId] = fCorr;
                b.tagV[fId] = tag;
                if (recMine) {
                    b.corrBy[fId] = c;
                    b.corrPos[fId] = recPosMine;
                    b.tagR[fId] = tag;
                }
            }
            // A cell whose first inner iteration finds no downwind face with room writes nothing, and will write nothing in
            // later sweeps either until a neighbour's correction changes its alpha: mark it dormant (see k_bound_run)
            if (!hadRoom && q == 0) oobState[c] = 3;
            (void)own;
            if (succMask) {
                __threadfence();  // corrections visible before any successor is released
                __syncwarp(M);
                bool got = false;
                if (valid && ((succMask >> q) & 1ull)) got = (atomicSub(&depLeft[other], 1) == 1);  // last predecessor: runs next here
                unsigned rel = __ballot_sync(M, got) & M;
                while (rel) {
                    const int k = __ffs(rel) - 1;
                    rel &= rel - 1;
                    const int y = __shfl_sync(M, other, k);
                    if (sp < SV_BSTACK) stack[sp++] = y; else if (q == 0) atomicOr(&ctl->err, SVERR_LIST);
                }
            }
            if (sp == 0) break;
            c = stack[--sp];
            i = oobIdx[c];
            __threadfence();  // acquire side of the release above
        }
    }
}

// sweep s, last step (advectionTemplates.C:164-192,207-208): apply each recorded correction once to
// alpha[own]/alpha[nei]/dVf, in the order of the reference's correctedFaces list (= ascending
// corrector cell, then position in its first-iteration face list); then build the next sweep's list.
template <int SV_MAXBF>
__global__ void __launch_bounds__(128) k_bound_apply(MeshDev m, Ctl* ctl, int s, const int* affList, const unsigned int* near1,
                                                     double* alpha, double* dVf, BoundScratch b, int* oobListNext,
                                                     unsigned char* oobState, const unsigned int* __restrict__ phiBits = nullptr)
{
    const int n = ctl->nAff[s];
    const int tag = boundTag(ctl, s);
    // the list sweep s consumed (deps and run are complete) becomes the empty output list of sweep s+1;
    // nobody in this kernel reads its counter
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->nOob[s & 1] = 0;
    int delta = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = affList[i];
        double a = alpha[c];
        const bool was = oobGlobal(a);
        const int c0 = m.cellOff[c];
        int nfc = m.cellOff[c + 1] - c0;
        if (nfc > SV_MAXBF) nfc = SV_MAXBF;
        int fl[SV_MAXBF];
        long long key[SV_MAXBF];
        double cvv[SV_MAXBF];
        unsigned long long ownM = 0;
        int nf = 0;
        for (int q0 = 0; q0 < nfc; q0 += 8) {
            int ff[8], tr[8], by[8], ps[8], ow[8];
            double cv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ff[j] = (q0 + j < nfc) ? m.cellFaces[c0 + q0 + j] : -1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                tr[j] = 0; by[j] = 0; ps[j] = 0; ow[j] = 0; cv[j] = 0.0;
                if (ff[j] >= 0) {
                    tr[j] = b.tagR[ff[j]];
                    by[j] = b.corrBy[ff[j]];
                    ps[j] = b.corrPos[ff[j]];
                    cv[j] = b.corr[ff[j]];
                    ow[j] = m.owner[ff[j]];
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (ff[j] < 0 || tr[j] != tag) continue;  // empty-patch faces are never recorded
                fl[nf] = ff[j];
                key[nf] = ((long long)by[j] << 20) | (long long)ps[j];
                cvv[nf] = cv[j];
                if (ow[j] == c) ownM |= 1ull << nf;
                nf++;
            }
        }
        for (int x = 1; x < nf; ++x) {  // insertion sort by key
            const long long kx = key[x];
            const int fx = fl[x];
            const double vx = cvv[x];
            const bool ox = (ownM >> x) & 1ull;
            int y = x - 1;
            while (y >= 0 && key[y] > kx) {
                key[y + 1] = key[y];
                fl[y + 1] = fl[y];
                cvv[y + 1] = cvv[y];
                ownM = (ownM & ~(1ull << (y + 1))) | (((ownM >> y) & 1ull) << (y + 1));
                --y;
            }
            key[y + 1] = kx;
            fl[y + 1] = fx;
            cvv[y + 1] = vx;
            ownM = (ownM & ~(1ull << (y + 1))) | ((unsigned long long)ox << (y + 1));
        }
        const double Vc = m.V[c];
        for (int x = 0; x < nf; ++x) {
            if ((ownM >> x) & 1ull) {
                a -= cvv[x] / Vc;
                dVf[fl[x]] = dVf[fl[x]] + cvv[x];  // setFaceValue(dVf_, facei, corrVf): once, by the owner
                // svof_step_host reads alphaPhi back on the faces its phi bitmap marks (a cell with liquid on one side): a
                // correction elsewhere (an empty cell overfilled at Courant > 1) makes it fall back to the full field
                if (phiBits && !((phiBits[fl[x] >> 5] >> (fl[x] & 31)) & 1u)) ctl->packUnsafe = 1;
            } else {
                a += cvv[x] / Vc;
            }
        }
        alpha[c] = a;
        delta += int(oobGlobal(a)) - int(was);
        if (oobBound(a) && bitTest(near1, c)) {
            if (!(oobState[c] == 3 && nf == 0)) {  // dormant cells (no room, alpha unchanged) stay off the next list
                oobState[c] = 1;
                oobListNext[atomicAdd(&ctl->nOob[(s + 1) & 1], 1)] = c;
            }
        } else {
            oobState[c] = 0;
        }
    }
    if (delta) atomicAdd(&ctl->nearOob[s + 1], delta);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->nearOob[s + 1], ctl->nearOob[s]);
}


// A12 for the near2 cells + alphaPhi on the faces they own + their bits of the next mixed bitmap
__global__ void __launch_bounds__(128) k_near_finalize(MeshDev m, const int* near2List, Ctl* ctl, double* alpha, const double* dVf,
                                                       double* alphaPhi, unsigned int* mixedNext, double dt, StepParams sp,
                                                       unsigned char* oobState)
{
    const int n = ctl->nNear2;
    double mn = SV_VGREAT, mx = -SV_VGREAT;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = near2List[i];
        const double a0 = alpha[c];
        mn = dmin(mn, a0);  // "After conservative bounding" (advectionTemplates.C:207-212), before snap/clip
        mx = dmax(mx, a0);
        const double a = snapClip(a0, sp.snapTol, sp.clip);
        alpha[c] = a;
        oobState[c] = 0;
        if ((sp.mixedTol < a) && (a < 1.0 - sp.mixedTol)) atomicOr(&mixedNext[c >> 5], 1u << (c & 31));
        for (int k = m.cellOff[c]; k < m.cellOff[c + 1]; ++k) {
            const int f = m.cellFaces[k];
            if (m.owner[f] == c && faceActive(m, f)) alphaPhi[f] = dVf[f] / dt;  // advectionTemplates.C:417
        }
    }
    blockMinMax(mn, mx, &ctl->minNearF, &ctl->maxNearF);
}

// volScalarField::correctBoundaryConditions for zeroGradient / fixedValue / inletOutlet -- thread per boundary face
struct PatchDev {
    int start, size, kind, bc;
    double value;
};
__global__ void k_alpha_bc(MeshDev m, const PatchDev* patches, const int* bPatch, const double* __restrict__ alpha,
                           const double* __restrict__ phi, double* alphaB)
{
    const int bf = blockIdx.x * blockDim.x + threadIdx.x;
    if (bf >= m.nBF) return;
    const PatchDev p = patches[bPatch[bf]];
    if (p.kind == 2) return;  // processor: filled by the halo exchange
    double v = 0.0;
    if (p.kind == 0) {
        const int f = m.nIF + bf;
        const double internal = alpha[m.owner[f]];
        if (p.bc == 1) v = p.value;
        else if (p.bc == 2) {
            const double vf = 1.0 - pos0(phi[f]);
            v = vf * p.value + (1.0 - vf) * internal;
        } else v = internal;
    }
    alphaB[bf] = v;
}

// ---- end-to-end (host buffers) transfer reduction --------------------------------------------------
// The PCIe link, not the GPU, bounds svof_step_host (816 MB in, 538 MB out per step at 256^3).  Two exact
// reductions of what has to cross it:
//  * U is only read by the interface-velocity interpolation (advection.C:91,126): the cells sharing a vertex
//    with a cut cell.  k_mark_u_cells lists them; the host gathers just those rows of U.
//  * alpha/alphaPhi come back as (index, value) deltas against what the host buffer already holds (the
//    previous step's result == alpha.oldTime on the device); most of a VOF domain does not change bitwise.
__global__ void k_mark_u_cells(MeshDev m, const int* mixedCells, const int* cellStatus, Ctl* ctl, unsigned int* uBits, int* uList,
                               int cap)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // cellStatus == nullptr: every interface cell (a superset of the cut ones) -- the zero-copy host step marks the rows
        // before the plane positioning has run, so that the pull overlaps it
        if (cellStatus && cellStatus[i] != 0) continue;
        const int c = mixedCells[i];
        for (int k = m.cellPtOff[c]; k < m.cellPtOff[c + 1]; ++k) {
            const int p = m.cellPts[k];
            for (int j = m.ptCellOff[p]; j < m.ptCellOff[p + 1]; ++j) {
                const int y = m.ptCells[j];
                const unsigned int bit = 1u << (y & 31);
                if (uBits[y >> 5] & bit) continue;
                const unsigned int old = atomicOr(&uBits[y >> 5], bit);
                if (!(old & bit)) {
                    const int pos = atomicAdd(&ctl->nUCells, 1);
                    if (pos < cap) uList[pos] = y;
                }
            }
        }
    }
}
__global__ void k_clear_u_bits(const int* uList, const Ctl* ctl, unsigned int* uBits, int cap)
{
    const int n = min(ctl->nUCells, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) uBits[uList[i] >> 5] = 0u;
}
__global__ void k_scatter_u(const int* uList, int n, const double* packed, double* U)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = uList[i];
    U[3 * (size_t)c] = packed[3 * (size_t)i];
    U[3 * (size_t)c + 1] = packed[3 * (size_t)i + 1];
    U[3 * (size_t)c + 2] = packed[3 * (size_t)i + 2];
}
//  * phi enters the step only as a factor of an alpha (upwind transport phi*alpha_up, geometric flux of cut cells, the
//    bounding of out-of-bounds cells, the downwind test of cut cells) or on a boundary face (inletOutlet patch values):
//    on an internal face with alpha == 0 exactly on BOTH sides its value cannot change any result (the product is a signed
//    zero, alphaPhi compares equal), so only the other faces have to cross PCIe.  k_phi_need_bits publishes them as a bitmap
//    over faces; the host gathers the marked entries of the caller's phi in face order and k_phi_scatter puts them back.
// blockCnt[1 + b] += marked faces of the 1024-face block b (zeroed by the caller; the host turns it into the prefix sum)
// tol = 0: exact (alpha != 0).  tol > 0 (option "sparse_phi_exp"): cells with |alpha| <= tol count as empty -- with snapTol 0 the
// support of alpha grows by one cell layer per step downstream (upwind transport of round-off-sized values), which this
// keeps out of the bitmap at the price of an O(tol) difference from the full-field call.
__global__ void k_phi_need_bits(MeshDev m, const double* __restrict__ alpha, unsigned int* bits, int nWordsF, int* blockCnt, double tol)
{
    const long long fl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool need = false;
    if (fl < m.nFaces) {
        const int f = (int)fl;
        need = (f >= m.nIF) || (fabs(__ldg(alpha + __ldg(m.owner + f))) > tol) || (fabs(__ldg(alpha + __ldg(m.neighbour + f))) > tol);
    }
    const unsigned int w = __ballot_sync(0xffffffffu, need);
    if ((threadIdx.x & 31) == 0 && (fl >> 5) < nWordsF) {
        bits[fl >> 5] = w;
        if (w) atomicAdd(blockCnt + 1 + (fl >> 10), __popc(w));
    }
}
// blockOff[b] = number of marked faces before bitmap word 32*b (host prefix sum); one thread per bitmap word
__global__ void k_phi_scatter(const unsigned int* __restrict__ bits, const int* __restrict__ blockOff, int nWordsF,
                              const double* __restrict__ packed, double* phi)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned int word = (w < nWordsF) ? bits[w] : 0u;
    const int cnt = __popc(word);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (!word) return;
    int pos = blockOff[w >> 5] + incl - cnt;
    while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        phi[((size_t)w << 5) + b] = packed[pos++];
    }
}
// ---- svof_step_host with PINNED caller buffers: the kernels read the caller's phi / U and write its alpha / alphaPhi
// directly over PCIe (zero copy), so the host neither gathers nor scatters and never waits inside the step.
// k_phi_pull = k_phi_need_bits + the pull of the marked entries: lane b of a warp handles face 32 w + b, so the reads of
// a fully marked word are one 256-byte request.
__global__ void k_phi_pull(MeshDev m, const double* __restrict__ alpha, double tol, const double* phiHost, double* phiDev,
                           unsigned int* bits, int nWordsF, Ctl* ctl)
{
    const long long fl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool need = false;
    if (fl < m.nFaces) {
        const int f = (int)fl;
        need = (f >= m.nIF) || (fabs(__ldg(alpha + __ldg(m.owner + f))) > tol) || (fabs(__ldg(alpha + __ldg(m.neighbour + f))) > tol);
        if (need) phiDev[f] = phiHost[f];
    }
    const unsigned int w = __ballot_sync(0xffffffffu, need);
    if ((threadIdx.x & 31) == 0 && (fl >> 5) < nWordsF) {
        bits[fl >> 5] = w;
        if (w) atomicAdd(&ctl->nPhiPulled, __popc(w));
    }
}
// Every face of a needBounding cell (near1: the interface cells and their face neighbours), marked or not: the bounding of an
// EMPTY near1 cell that went out of bounds reads phi on all its faces, and a PCIe read issued from inside the sweep would queue
// behind the result pushes that share the link.  Entries that were not marked yet are pulled, stored and marked.
__global__ void k_phi_pull_near(MeshDev m, const int* __restrict__ near2List, const unsigned int* __restrict__ near1, Ctl* ctl,
                                const double* phiHost, double* phiDev, unsigned int* bits)
{
    const int n = ctl->nNear2;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = near2List[i];
        if (!bitTest(near1, c)) continue;
        for (int k = m.cellOff[c]; k < m.cellOff[c + 1]; ++k) {
            const int f = m.cellFaces[k];
            const unsigned int bit = 1u << (f & 31);
            if (__ldcg(bits + (f >> 5)) & bit) continue;
            phiDev[f] = phiHost[f];   // both cells of a face may do this: same value
            __threadfence();
            if (!(atomicOr(bits + (f >> 5), bit) & bit)) cnt++;
        }
    }
    if (cnt) atomicAdd(&ctl->nPhiPulled, cnt);
}
// rows of U next to cut cells (bitmap from k_mark_u_cells), straight from the caller's buffer
__global__ void k_u_pull(const unsigned int* __restrict__ uBits, int nCells, const double* UHost, double* UDev)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells || !((uBits[c >> 5] >> (c & 31)) & 1u)) return;
    UDev[3 * (size_t)c] = UHost[3 * (size_t)c];
    UDev[3 * (size_t)c + 1] = UHost[3 * (size_t)c + 1];
    UDev[3 * (size_t)c + 2] = UHost[3 * (size_t)c + 2];
}
// alpha cells whose bit pattern changed go straight into the caller's buffer (which holds the previous field)
__global__ void k_alpha_push(const double* __restrict__ cur, const double* __restrict__ ref, int n, double* host, Ctl* ctl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool changed = false;
    if (i < n) {
        const double v = cur[i];
        changed = __double_as_longlong(v) != __double_as_longlong(ref[i]);
        if (changed) host[i] = v;
    }
    const unsigned int mk = __ballot_sync(0xffffffffu, changed);
    if ((threadIdx.x & 31) == 0 && mk) atomicAdd(&ctl->nDeltaA, __popc(mk));
}
// alphaPhi can only be non-zero on the faces this step's bitmap marks: those are written; faces the previous step marked
// and this one does not are zero now; if a bounding correction landed elsewhere (ctl->packUnsafe) every face is written
__global__ void k_alphaphi_push(const unsigned int* __restrict__ bitsCur, const unsigned int* __restrict__ bitsPrev,
                                const double* __restrict__ alphaPhi, double* host, int nFaces, Ctl* ctl)
{
    const long long fl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (fl >= nFaces) return;
    const int w = (int)(fl >> 5);
    const unsigned int bit = 1u << lane;
    const bool all = ctl->packUnsafe != 0;
    const unsigned int cur = bitsCur[w], prev = bitsPrev[w];
    if (all || (cur & bit)) host[fl] = alphaPhi[fl];
    else if (prev & bit) host[fl] = 0.0;
    if (lane == 0) {
        const int cnt = all ? 32 : __popc(cur | prev);
        if (cnt) atomicAdd(&ctl->nDeltaF, cnt);
    }
}
// ---- the same pushes split by WHEN a value is final.  A cell outside near2 and a face whose owner is outside near2 get their
// final alpha / alphaPhi from the streaming kernel (snap/clip included), which runs beside the interface kernels: those values
// ("early") cross PCIe on the streaming kernel's stream while the interface chain is still working.  near2 cells and the faces
// they own are final after k_near_finalize ("late"): a list-driven kernel at the end of the step.
__global__ void k_alpha_push_early(const double* __restrict__ cur, const double* __restrict__ ref, const unsigned int* __restrict__ near2,
                                   int n, double* host, Ctl* ctl)
{
    // a small persistent grid (the launcher gives it a fraction of every SM): the interface kernels on the other stream must
    // keep finding free CTA slots while these warps sit on PCIe back-pressure
    const int lane = threadIdx.x & 31;
    const long long nWarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nW = (n + 31) >> 5;
    int cnt = 0;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nW; w += nWarps) {
        const int i = (int)(w << 5) + lane;
        bool changed = false;
        if (i < n && !((near2[w] >> lane) & 1u)) {
            const double v = cur[i];
            changed = __double_as_longlong(v) != __double_as_longlong(ref[i]);
            if (changed) host[i] = v;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, changed));
    }
    if (lane == 0 && cnt) atomicAdd(&ctl->nDeltaA, cnt);
}
__global__ void k_alpha_push_late(const int* __restrict__ near2List, Ctl* ctl, const double* __restrict__ cur,
                                  const double* __restrict__ ref, double* host)
{
    const int n = ctl->nNear2;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = near2List[i];
        const double v = cur[c];
        if (__double_as_longlong(v) != __double_as_longlong(ref[c])) {
            host[c] = v;
            cnt++;
        }
    }
    if (cnt) atomicAdd(&ctl->nDeltaA, cnt);
}
__global__ void k_alphaphi_push_early(MeshDev m, const unsigned int* __restrict__ bitsCur, const unsigned int* __restrict__ bitsPrev,
                                      const unsigned int* __restrict__ near2, const double* __restrict__ alphaPhi, double* host,
                                      int nFaces, Ctl* ctl)
{
    const int lane = threadIdx.x & 31;
    const long long nWarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nW = (nFaces + 31) >> 5;
    int cnt = 0;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nW; w += nWarps) {
        const unsigned int cur = bitsCur[w], prev = bitsPrev[w];
        if (!(cur | prev)) continue;   // warp-uniform
        const long long fl = (w << 5) + lane;
        const unsigned int bit = 1u << lane;
        bool wrote = false;
        if (fl < nFaces && ((cur | prev) & bit) && !bitTest(near2, __ldg(m.owner + fl))) {
            host[fl] = (cur & bit) ? alphaPhi[fl] : 0.0;
            wrote = true;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, wrote));
    }
    if (lane == 0 && cnt) atomicAdd(&ctl->nDeltaF, cnt);
}
__global__ void k_alphaphi_push_late(MeshDev m, const int* __restrict__ near2List, Ctl* ctl, const unsigned int* __restrict__ bitsCur,
                                     const unsigned int* __restrict__ bitsPrev, const double* __restrict__ alphaPhi, double* host)
{
    const int n = ctl->nNear2;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = near2List[i];
        for (int k = m.cellOff[c]; k < m.cellOff[c + 1]; ++k) {
            const int f = m.cellFaces[k];
            if (m.owner[f] != c) continue;
            const unsigned int bit = 1u << (f & 31);
            const unsigned int cur = bitsCur[f >> 5], prev = bitsPrev[f >> 5];
            if ((cur | prev) & bit) {
                host[f] = (cur & bit) ? alphaPhi[f] : 0.0;
                cnt++;
            }
        }
    }
    if (cnt) atomicAdd(&ctl->nDeltaF, cnt);
}
// a bounding correction landed on a face outside the bitmap (ctl->packUnsafe, rare): every face is written
__global__ void k_alphaphi_push_all(const double* __restrict__ alphaPhi, double* host, int nFaces, Ctl* ctl)
{
    if (!ctl->packUnsafe) return;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < nFaces; f += (long long)gridDim.x * blockDim.x)
        host[f] = alphaPhi[f];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->nDeltaF, nFaces);
}
// the reverse of k_phi_scatter: packed[k] = field[f] for the marked faces in face order (alphaPhi read-back: alphaPhi can only
// be non-zero on a face whose upwind cell held liquid, i.e. on the faces the step's phi bitmap marks)
__global__ void k_phi_pack(const unsigned int* __restrict__ bits, const int* __restrict__ blockOff, int nWordsF,
                           const double* __restrict__ field, double* __restrict__ packed)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned int word = (w < nWordsF) ? bits[w] : 0u;
    const int cnt = __popc(word);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (!word) return;
    int pos = blockOff[w >> 5] + incl - cnt;
    while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        packed[pos++] = field[((size_t)w << 5) + b];
    }
}
// entries whose bit pattern changed: (index, value) appended with warp-aggregated atomics; if prev != nullptr
// it is brought up to date at the same time
__global__ void k_delta(const double* __restrict__ cur, double* prev, const double* __restrict__ ref, long long n, int* counter,
                        int* idx, double* val, int cap)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool changed = false;
    double v = 0.0;
    if (i < n) {
        v = cur[i];
        const double o = ref ? ref[i] : prev[i];
        changed = __double_as_longlong(v) != __double_as_longlong(o);
        if (changed && prev) prev[i] = v;
    }
    const unsigned int mask = __ballot_sync(0xffffffffu, changed);
    if (mask) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (changed) {
            const int pos = base + __popc(mask & ((1u << lane) - 1u));
            if (pos < cap) {
                idx[pos] = (int)i;
                val[pos] = v;
            }
        }
    }
}

// ---- on-demand outputs ---------------------------------------------------------------------
// dVf_ as the reference holds it after advect(): alphaPhi*dt is NOT bitwise dVf, so it is rebuilt:
// upwind transport (from alpha.oldTime) everywhere, the stored scratch on faces of near2 cells.
__global__ void k_materialize_dvf(MeshDev m, const double* aOld, const double* alphaB, const double* phi, double dt,
                                  const unsigned int* near2, const double* dVfScratch, double* out)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.nFaces) return;
    const int own = m.owner[f];
    const int nei = (f < m.nIF) ? m.neighbour[f] : -1;
    if (!faceActive(m, f)) {
        out[f] = 0.0;
        return;
    }
    if (bitTest(near2, own) || (nei >= 0 && bitTest(near2, nei))) {
        out[f] = dVfScratch[f];
        return;
    }
    double dvf = 0.0;
    upwindDVf(m, own, f, false, (nei >= 0) ? nei : (-1 - (f - m.nIF)), phi[f], aOld, alphaB, dt, dvf);
    out[f] = dvf;
}

// deterministic sum(alpha*V): fixed-shape tree (block partials, then one block)
__global__ void k_volume_partial(const double* alpha, const double* V, int n, double* partial)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += alpha[i] * V[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

#endif  // !SV_VARIANT

}  // namespace svof
