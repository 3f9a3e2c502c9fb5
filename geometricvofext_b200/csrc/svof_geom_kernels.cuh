// svof_geom_kernels.cuh -- the kernels that depend on the compiled polyhedron capacity variant
// (Caps<MAXFV, MAXCF, MAXCP>).  Each variant is instantiated in its own translation unit
// (svof_inst.cu compiled with -DSV_VARIANT=n) so the variants build in parallel.
#pragma once
#include "svof_kernels.cuh"
#include "svof_plic_group.cuh"
#include "svof_plic_warp.cuh"

namespace svof {

typedef Caps<4, 6, 8> CapsHex;         // hexahedra (blockMesh), flat faces
typedef Caps<8, 16, 32> CapsSmall;     // tets/prisms/small polyhedra
typedef Caps<16, 40, 72> CapsPoly;     // polyDualMesh cells (~14 faces, 24+ vertices)
typedef Caps<16, 200, 128> CapsSplit;  // splitWarpedFace local triangulations
typedef Caps<4, 24, 16> CapsHexSplit;  // hexahedra with splitWarpedFace: 6 x 4 triangles, 8 + 6 points (variant index 4)

// host-callable launchers, one explicit instantiation per variant
template <class CP>
struct GeoLaunch {
    static void plic(cudaStream_t st, int grid, MeshDev m, const int* mixedCells, Ctl* ctl, const double* alpha, const double* iN,
                     int split, int* cellStatus, double* iD, double* iC, double* iS);
    static void faceFlux(cudaStream_t st, int grid, MeshDev m, const int2* work, Ctl* ctl, const int* mixedCells, const double* iN,
                         const double* iD, const double* Un0, const double* phi, double dt, double* dVfGeo);
    static void cutFaces(cudaStream_t st, int nPolys, int nVerts, const double* pts, const double* normals, const double* dists,
                         int* status, double* centres, double* areas, int* errOut);
    static void cutCells(cudaStream_t st, MeshDev m, int n, const int* cells, const double* normals, const double* dists, int* status,
                         double* vof, double* subVol, double* ic, double* ia, int* errOut);
    static void findDistance(cudaStream_t st, MeshDev m, int n, const int* cells, const double* alphas, const double* normals,
                             int split, int* status, double* dists, double* ic, double* ia, int* errOut);
    static void faceFluxes(cudaStream_t st, MeshDev m, int n, const int* faces, const double* normals, const double* dists,
                           const double* Un0, double dt, const double* phi, double* out, int* errOut);
    // reconstruction::interface(): up to maxPolyPoints() points per mixed cell into polyPts[i*maxPolyPoints()..], count in polyCount[i]
    static int maxPolyPoints() { return CP::MAXEP; }
    static void plicPolygons(cudaStream_t st, int grid, MeshDev m, const int* mixedCells, Ctl* ctl, const double* iN, const double* iD,
                             double* polyPts, int* polyCount);
    // reconstruction::subCellFaces(): raw polygons of the submerged sub-cell per mixed cell (fixed strides), merged on the host
    static void subCellFaces(cudaStream_t st, int grid, MeshDev m, const int* mixedCells, Ctl* ctl, const double* iN, const double* iD,
                             int maxFaces, int maxPts, double* pts, int* faceSize, int* nFaces, double* centre);
    // reconstruction::mapAlphaField: alpha <- calcSubCell(cell, interfaceN, interfaceD).VOF where lower <= alpha <= upper
    static void mapAlpha(cudaStream_t st, MeshDev m, const double* iN, const double* iD, double lower, double upper, double* alpha, Ctl* ctl);
};

#ifdef SV_VARIANT
// A3-A5: plane positioning, thread per mixed cell
template <class CP>
__global__ void __launch_bounds__(128) k_plic(MeshDev m, const int* mixedCells, Ctl* ctl, const double* __restrict__ alpha,
                                              const double* iN, int split, int* cellStatus, double* iD, double* iC, double* iS)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = mixedCells[i];
        int err = 0;
        PlicOut po;
        signedDistance<CP>(m, c, alpha[c], ld3(iN, c), split != 0, po, err);
        cellStatus[i] = po.status;
        if (po.wrote) {
            iD[c] = po.D;
            st3(iC, c, po.C);
            st3(iS, c, po.S);
        }
        if (err) atomicOr(&ctl->err, err);
    }
}

// A8+A9: thread per (cut cell, downwind face)
// (Round 1 tried staging the face in shared memory and an area-only streaming clip for the up to 11 calcSubFace
// evaluations of the Simpson rule, as in k_plic_group: 68 us against 66 us for this form -- not where the time goes.)
template <class CP>
__global__ void __launch_bounds__(128) k_face_flux(MeshDev m, const int2* work, Ctl* ctl, const int* mixedCells, const double* iN,
                                                   const double* iD, const double* Un0, const double* __restrict__ phi, double dt,
                                                   double* dVfGeo)
{
    int n = ctl->nWork;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int2 w = work[i];
        const int c = mixedCells[w.x], f = w.y;
        int err = 0;
        dVfGeo[f] = faceFlux<CP>(m, f, ld3(iN, c), iD[c], Un0[w.x], dt, phi[f], m.magSf[f], err);
        if (err) atomicOr(&ctl->err, err);
    }
}

// ---- geometry primitives exposed through the C ABI (unit-test surface) -------------------------
template <class CP>
__global__ void k_cut_faces(int nPolys, int nVerts, const double* pts, const double* normals, const double* dists, int* status,
                            double* centres, double* areas, int* errOut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nPolys) return;
    d3 fp[CP::MAXFV], c, a, ip[CP::MAXIP];
    int nip, err = 0;
    const int nv = nVerts > CP::MAXFV ? CP::MAXFV : nVerts;
    for (int k = 0; k < nv; ++k) fp[k] = ld3(pts, (int64_t)i * nVerts + k);
    status[i] = clipFace<CP>(fp, nv, ld3(normals, i), dists[i], c, a, ip, nip, err);
    st3(centres, i, c);
    st3(areas, i, a);
    if (err || nVerts > CP::MAXFV) atomicOr(errOut, err | (nVerts > CP::MAXFV ? SVERR_FACE_VERTS : 0));
}
template <class CP>
__global__ void k_cut_cells(MeshDev m, int n, const int* cells, const double* normals, const double* dists, int* status, double* vof,
                            double* subVol, double* ic, double* ia, int* errOut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SubCellOut sc;
    int err = 0;
    subCell<CP>(m, cells[i], ld3(normals, i), dists[i], false, sc, err);
    status[i] = sc.status;
    vof[i] = sc.VOF;
    subVol[i] = sc.subVol;
    st3(ic, i, sc.iC);
    st3(ia, i, sc.iS);
    if (err) atomicOr(errOut, err);
}
// reconstruction::interface() (reconstruction.C:787-835) + cutCell::interfacePoints (cutCell.C:545-608): thread per mixed
// cell; the cut is re-evaluated WITHOUT splitWarpedFace (as the reference does), the interface edge points are sorted
// by angle about the interface centre in the plane of the interface (stable, ascending) and points within 1e-8 rad of
// their predecessor are dropped.
// cutCell::interfacePoints (cutCell.C:545-608): the interface edge points sorted by angle about the interface centre in the
// plane of the interface (stable, ascending); points within 1e-8 rad of their predecessor are dropped.  Returns the count.
template <class CP>
__device__ __forceinline__ int interfacePolygonDev(const SubCellOut& sc, const d3* ep, double* out)
{
    int cnt = 0;
    const d3 zhat = sc.iS / mag(sc.iS);
    d3 xhat = ep[0] - sc.iC;
    xhat = xhat - dot(xhat, zhat) * zhat;
    xhat /= mag(xhat);
    d3 yhat = cross(zhat, xhat);
    yhat /= mag(yhat);
    double ang[CP::MAXEP];
    short ord[CP::MAXEP];
    for (int q = 0; q < sc.nEp; ++q) {
        const d3 d = ep[q] - sc.iC;
        const double a = atan2(dot(d, yhat), dot(d, xhat));
        int j = q - 1;  // stable insertion: equal angles keep their original order
        while (j >= 0 && ang[j] > a) {
            ang[j + 1] = ang[j];
            ord[j + 1] = ord[j];
            --j;
        }
        ang[j + 1] = a;
        ord[j + 1] = (short)q;
    }
    for (int pi = 0; pi < sc.nEp; ++pi) {
        if (pi > 0 && !(fabs(ang[pi] - ang[pi - 1]) > 1e-8)) continue;
        st3(out, cnt, ep[ord[pi]]);
        cnt++;
    }
    return cnt;
}

template <class CP>
__global__ void __launch_bounds__(128) k_plic_polygons(MeshDev m, const int* mixedCells, Ctl* ctl, const double* iN, const double* iD,
                                                       double* polyPts, int* polyCount)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = mixedCells[i];
        d3 ep[CP::MAXEP];
        SubCellOut sc;
        sc.epOut = ep;
        int err = 0, cnt = 0;
        subCell<CP>(m, c, ld3(iN, c), iD[c], false, sc, err);
        if (sc.status == 0 && sc.nEp > 0) cnt = interfacePolygonDev<CP>(sc, ep, polyPts + 3 * (size_t)i * CP::MAXEP);
        polyCount[i] = cnt;
        if (err) atomicOr(&ctl->err, err);
    }
}

// reconstruction::subCellFaces() (reconstruction.C:838-891) + the collecting half of cutCell::calcSubCell (cutCell.C:443-475):
// thread per mixed cell; per cut cell the clipped polygon of every cut face, every fully submerged face as it is, then the
// interface polygon, into fixed-stride buffers (faceSize[i*maxFaces + k] points each, consecutive in pts[i*maxPts ...]).
// The point merge and the orientation fix of updateSubCellPointsandFaces (cutCell.C:239-290) run on the host.
template <class CP>
__global__ void __launch_bounds__(128) k_subcell_faces(MeshDev m, const int* mixedCells, Ctl* ctl, const double* iN, const double* iD,
                                                       int maxFaces, int maxPts, double* pts, int* faceSize, int* nFaces, double* centre)
{
    const int n = ctl->nMixed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = mixedCells[i];
        const d3 nn = ld3(iN, c);
        const double D = iD[c];
        d3 ep[CP::MAXEP];
        SubCellOut sc;
        sc.epOut = ep;
        sc.wantCentre = true;
        int err = 0, nf = 0, np = 0;
        subCell<CP>(m, c, nn, D, false, sc, err);
        if (sc.status == 0) {
            double* out = pts + 3 * (size_t)i * maxPts;
            int* fs = faceSize + (size_t)i * maxFaces;
            const int c0 = __ldg(m.cellOff + c), c1 = __ldg(m.cellOff + c + 1);
            for (int k = c0; k < c1; ++k) {
                d3 fp[CP::MAXFV], sp[2 * CP::MAXFV], fc, fa, ip[CP::MAXIP];
                int nip, nsp = 0;
                const int nv = loadFace<CP>(m, __ldg(m.cellFaces + k), fp, err);
                const int st = clipFace<CP>(fp, nv, nn, D, fc, fa, ip, nip, err, sp, &nsp);
                if (st > 0) continue;
                const d3* src = (st == 0) ? sp : fp;
                const int cnt = (st == 0) ? nsp : nv;
                if (nf >= maxFaces || np + cnt > maxPts) { err |= SVERR_CELL_FACES; break; }
                for (int q = 0; q < cnt; ++q) st3(out, np + q, src[q]);
                fs[nf++] = cnt;
                np += cnt;
            }
            if (sc.nEp > 0 && nf < maxFaces && np + sc.nEp <= maxPts) {
                const int cnt = interfacePolygonDev<CP>(sc, ep, out + 3 * (size_t)np);
                if (cnt > 0) {
                    fs[nf++] = cnt;
                    np += cnt;
                }
            }
            st3(centre, i, sc.subCentre);
        }
        nFaces[i] = nf;
        if (err) atomicOr(&ctl->err, err);
    }
}

// reconstruction::mapAlphaField (reconstruction.C:751-768): thread per cell, the cut is evaluated WITHOUT splitWarpedFace
template <class CP>
__global__ void __launch_bounds__(128) k_map_alpha(MeshDev m, const double* iN, const double* iD, double lower, double upper, double* alpha,
                                                   Ctl* ctl)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.nCells) return;
    const double a = alpha[c];
    if (!(a >= lower && a <= upper)) return;
    SubCellOut sc;
    int err = 0;
    subCell<CP>(m, c, ld3(iN, c), iD[c], false, sc, err);
    alpha[c] = sc.VOF;
    if (err) atomicOr(&ctl->err, err);
}
template <class CP>
__global__ void k_find_distance(MeshDev m, int n, const int* cells, const double* alphas, const double* normals, int split,
                                int* status, double* dists, double* ic, double* ia, int* errOut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PlicOut po;
    int err = 0;
    signedDistance<CP>(m, cells[i], alphas[i], ld3(normals, i), split != 0, po, err);
    status[i] = po.status;
    dists[i] = po.D;
    st3(ic, i, po.C);
    st3(ia, i, po.S);
    if (err) atomicOr(errOut, err);
}
template <class CP>
__global__ void k_face_fluxes(MeshDev m, int n, const int* faces, const double* normals, const double* dists, const double* Un0,
                              double dt, const double* phi, double* out, int* errOut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int err = 0;
    const int f = faces[i];
    out[i] = faceFlux<CP>(m, f, ld3(normals, i), dists[i], Un0[i], dt, phi[i], m.magSf[f], err);
    if (err) atomicOr(errOut, err);
}


template <class CP>
void GeoLaunch<CP>::plic(cudaStream_t st, int grid, MeshDev m, const int* mixedCells, Ctl* ctl, const double* alpha, const double* iN,
                         int split, int* cellStatus, double* iD, double* iC, double* iS)
{
    // warp-cooperative kernel: self-contained groups of 8 lanes per cell, 4 cells per warp, staging sized from the mesh's maxima
    static const bool useGroup = getenv("SVOF_PLIC") && !strcmp(getenv("SVOF_PLIC"), "group");  // round-1 CTA-phase kernel (A/B runs)
    const PlicWarpLayout L = PlicWarpLayout::make(m.maxLocalFaces, m.maxFV, m.maxLocalPts);
    const size_t cellBytes = (size_t)L.strideD * sizeof(double);
    if (!useGroup && 4 * cellBytes <= 200 * 1024) {
        int threads = SV_PW_THREADS;
        while (threads > 32 && cellBytes * (threads / SV_G) > 56 * 1024) threads >>= 1;
        const size_t smem = cellBytes * (threads / SV_G);
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(k_plic_warp<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
            configured = true;
        }
        // persistent: as many CTAs as fit on the device at once, or `grid` per SM when the caller caps it (> 0: the
        // "plic_ctas" option leaves room for the streaming kernel on every SM)
        static int resident = 0, residentThreads = 0, residentCap = -1;
        static size_t residentSmem = 0;
        if (!resident || residentThreads != threads || residentSmem != smem || residentCap != grid) {
            int perSm = 0, dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_plic_warp<CP>, threads, smem);
            if (grid > 0 && grid < perSm) perSm = grid;
            resident = (perSm > 0 ? perSm : 1) * (sms > 0 ? sms : 148);
            residentThreads = threads;
            residentSmem = smem;
            residentCap = grid;
        }
        k_plic_warp<CP><<<resident, threads, smem, st>>>(m, L, mixedCells, ctl, alpha, iN, split, cellStatus, iD, iC, iS);
        return;
    }
    // round-1 kernel: CTA of cpb cells x 8 lanes with leader phases (kept for A/B timing)
    size_t perCell = sizeof(GCellShared<CP>) + SV_G * LanePriv<CP>::DOUBLES * sizeof(double);
    if (CP::MAXCF > SV_G)   // per-face records for cells with more than one face per lane, sized from the mesh's maxima
        perCell += sizeof(double) * ((size_t)m.maxLocalFaces * FaceRec::doubles(m.maxFV) + (m.maxLocalFaces + 1) / 2);
    int threads = SV_PLIC_THREADS;
    while (threads > 32 && perCell * (threads / SV_G) > 100 * 1024) threads >>= 1;
    const size_t smem = perCell * (threads / SV_G);
    if (smem <= 200 * 1024) {
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(k_plic_group<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
            configured = true;
        }
        static int resident = 0, residentThreads = 0;
        static size_t residentSmem = 0;
        if (!resident || residentThreads != threads || residentSmem != smem) {
            int perSm = 0, dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_plic_group<CP>, threads, smem);
            resident = (perSm > 0 ? perSm : 1) * (sms > 0 ? sms : 148);
            residentThreads = threads;
            residentSmem = smem;
        }
        k_plic_group<CP><<<resident, threads, smem, st>>>(m, mixedCells, ctl, alpha, iN, split, cellStatus, iD, iC, iS);
    } else {
        k_plic<CP><<<148 * 8, 128, 0, st>>>(m, mixedCells, ctl, alpha, iN, split, cellStatus, iD, iC, iS);  // grid-stride, thread per cell
    }
}
template <class CP>
void GeoLaunch<CP>::faceFlux(cudaStream_t st, int grid, MeshDev m, const int2* work, Ctl* ctl, const int* mixedCells, const double* iN,
                             const double* iD, const double* Un0, const double* phi, double dt, double* dVfGeo)
{
    k_face_flux<CP><<<grid, 128, 0, st>>>(m, work, ctl, mixedCells, iN, iD, Un0, phi, dt, dVfGeo);
}
template <class CP>
void GeoLaunch<CP>::cutFaces(cudaStream_t st, int nPolys, int nVerts, const double* pts, const double* normals, const double* dists,
                             int* status, double* centres, double* areas, int* errOut)
{
    k_cut_faces<CP><<<(nPolys + 127) / 128, 128, 0, st>>>(nPolys, nVerts, pts, normals, dists, status, centres, areas, errOut);
}
template <class CP>
void GeoLaunch<CP>::cutCells(cudaStream_t st, MeshDev m, int n, const int* cells, const double* normals, const double* dists,
                             int* status, double* vof, double* subVol, double* ic, double* ia, int* errOut)
{
    k_cut_cells<CP><<<(n + 127) / 128, 128, 0, st>>>(m, n, cells, normals, dists, status, vof, subVol, ic, ia, errOut);
}
template <class CP>
void GeoLaunch<CP>::plicPolygons(cudaStream_t st, int grid, MeshDev m, const int* mixedCells, Ctl* ctl, const double* iN, const double* iD,
                                 double* polyPts, int* polyCount)
{
    k_plic_polygons<CP><<<grid, 128, 0, st>>>(m, mixedCells, ctl, iN, iD, polyPts, polyCount);
}
template <class CP>
void GeoLaunch<CP>::subCellFaces(cudaStream_t st, int grid, MeshDev m, const int* mixedCells, Ctl* ctl, const double* iN, const double* iD,
                                 int maxFaces, int maxPts, double* pts, int* faceSize, int* nFaces, double* centre)
{
    k_subcell_faces<CP><<<grid, 128, 0, st>>>(m, mixedCells, ctl, iN, iD, maxFaces, maxPts, pts, faceSize, nFaces, centre);
}
template <class CP>
void GeoLaunch<CP>::mapAlpha(cudaStream_t st, MeshDev m, const double* iN, const double* iD, double lower, double upper, double* alpha, Ctl* ctl)
{
    k_map_alpha<CP><<<(m.nCells + 127) / 128, 128, 0, st>>>(m, iN, iD, lower, upper, alpha, ctl);
}
template <class CP>
void GeoLaunch<CP>::findDistance(cudaStream_t st, MeshDev m, int n, const int* cells, const double* alphas, const double* normals,
                                 int split, int* status, double* dists, double* ic, double* ia, int* errOut)
{
    k_find_distance<CP><<<(n + 127) / 128, 128, 0, st>>>(m, n, cells, alphas, normals, split, status, dists, ic, ia, errOut);
}
template <class CP>
void GeoLaunch<CP>::faceFluxes(cudaStream_t st, MeshDev m, int n, const int* faces, const double* normals, const double* dists,
                               const double* Un0, double dt, const double* phi, double* out, int* errOut)
{
    k_face_fluxes<CP><<<(n + 127) / 128, 128, 0, st>>>(m, n, faces, normals, dists, Un0, dt, phi, out, errOut);
}
#endif  // SV_VARIANT

}  // namespace svof
