// svof_decomp.cpp -- host side of decomposed runs: cell partitioning and sub-domain extraction with ghost layers.
//
// What it replaces.  The reference runs in parallel the OpenFOAM way: decomposePar (scotch) writes one polyMesh per
// rank with processor patches, and every step exchanges alpha / stencil values / dVf across them ~10 times
// (zoneDistribute in reconstruction.C:97-107, syncProcPatches in advection.C:311-393 called 1 + 2 x sweeps times,
// gMin/gMax reductions in advectionTemplates.C:146-198).  Here each rank instead gets its owned cells PLUS `layers`
// point-neighbour layers of ghost cells, cut out of the global mesh with an order-preserving renumbering, and runs the
// unchanged single-domain step on that extended sub-mesh; ONE exchange per step refreshes ghost alpha from the owners
// (svof_halo_*).  Because local cell, face and point labels are monotone in the global ones, every order-dependent
// choice of the algorithm (ascending mixed-cell list, cells() face order, LS stencil order, ascending bounding sweep)
// is the single-domain one, so owned cells reproduce the single-domain result as long as the dependency radius of a
// step (LS stencil 1 + upwind plane 1 + one layer per bounding sweep) stays inside the ghost layers.
//
// Nothing here touches CUDA: the functions are usable on a CPU-only box (tests, pre-processing).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/svof.h"

struct svof_submesh {
    std::vector<double> points;
    std::vector<int32_t> faceOff, facePts, owner, neighbour;
    std::vector<svof_patch> patches;
    std::vector<int32_t> cellGlobal, cellOwnerRank, cellLayer, faceGlobal, pointGlobal, ownedLocal, faceOwnerRank, faceFlip;
    int32_t nCells = 0, nOwned = 0, nInternal = 0;
    std::string err;
};

namespace {
thread_local std::string g_decompError;

// weighted recursive coordinate bisection: split [begin,end) of `ids` into parts [p0, p0+np) along the longest axis
void rcb(std::vector<int32_t>& ids, size_t begin, size_t end, int p0, int np, const double* cx, const double* w, int32_t* out)
{
    if (np == 1) {
        for (size_t i = begin; i < end; ++i) out[ids[i]] = p0;
        return;
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = begin; i < end; ++i)
        for (int d = 0; d < 3; ++d) {
            const double v = cx[3 * (size_t)ids[i] + d];
            lo[d] = std::min(lo[d], v);
            hi[d] = std::max(hi[d], v);
        }
    int ax = 0;
    for (int d = 1; d < 3; ++d)
        if (hi[d] - lo[d] > (hi[ax] - lo[ax]) * (1.0 + 1e-12)) ax = d;
    // ties in the coordinate are broken by the cell label, so the result is deterministic
    std::sort(ids.begin() + begin, ids.begin() + end, [&](int32_t a, int32_t b) {
        const double va = cx[3 * (size_t)a + ax], vb = cx[3 * (size_t)b + ax];
        return va < vb || (va == vb && a < b);
    });
    const int npL = np / 2;
    double total = 0;
    for (size_t i = begin; i < end; ++i) total += w ? w[ids[i]] : 1.0;
    const double target = total * (double)npL / (double)np;
    double acc = 0;
    size_t cut = begin;
    while (cut < end && acc + 0.5 * (w ? w[ids[cut]] : 1.0) < target) {
        acc += w ? w[ids[cut]] : 1.0;
        ++cut;
    }
    // keep every part non-empty
    cut = std::max(cut, begin + (size_t)npL);
    cut = std::min(cut, end - (size_t)(np - npL));
    rcb(ids, begin, cut, p0, npL, cx, w, out);
    rcb(ids, cut, end, p0 + npL, np - npL, cx, w, out);
}
}  // namespace

extern "C" {

const char* svof_decomp_last_error(void) { return g_decompError.c_str(); }

int svof_partition_rcb(const svof_mesh* m, const double* cell_weight, int32_t n_parts, int32_t* cell_rank_out)
{
    if (!m || !cell_rank_out || n_parts < 1 || !m->points || !m->face_offsets || !m->face_points || !m->owner) {
        g_decompError = "svof_partition_rcb: null argument";
        return SVOF_ERR_INVALID_ARG;
    }
    const int nC = m->n_cells, nF = m->n_faces, nIF = m->n_internal_faces;
    if (n_parts > nC) {
        g_decompError = "svof_partition_rcb: more parts than cells";
        return SVOF_ERR_INVALID_ARG;
    }
    // cell "centres" for the bisection: mean of the vertex means of the cell's faces (partitioning only, not geometry)
    std::vector<double> cx((size_t)3 * nC, 0.0);
    std::vector<int> cnt(nC, 0);
    for (int f = 0; f < nF; ++f) {
        double s[3] = {0, 0, 0};
        const int a = m->face_offsets[f], b = m->face_offsets[f + 1];
        for (int k = a; k < b; ++k)
            for (int d = 0; d < 3; ++d) s[d] += m->points[3 * (size_t)m->face_points[k] + d];
        for (int d = 0; d < 3; ++d) s[d] /= (double)(b - a);
        const int o = m->owner[f];
        for (int d = 0; d < 3; ++d) cx[3 * (size_t)o + d] += s[d];
        cnt[o]++;
        if (f < nIF) {
            const int n = m->neighbour[f];
            for (int d = 0; d < 3; ++d) cx[3 * (size_t)n + d] += s[d];
            cnt[n]++;
        }
    }
    for (int c = 0; c < nC; ++c)
        for (int d = 0; d < 3; ++d) cx[3 * (size_t)c + d] /= (double)std::max(cnt[c], 1);
    std::vector<int32_t> ids(nC);
    std::iota(ids.begin(), ids.end(), 0);
    rcb(ids, 0, (size_t)nC, 0, n_parts, cx.data(), cell_weight, cell_rank_out);
    return SVOF_OK;
}

int svof_decompose(const svof_mesh* g, const int32_t* cell_rank, int32_t rank, int32_t layers, svof_submesh** out)
{
    if (!g || !cell_rank || !out || layers < 0) {
        g_decompError = "svof_decompose: null argument";
        return SVOF_ERR_INVALID_ARG;
    }
    try {
        const int nC = g->n_cells, nF = g->n_faces, nIF = g->n_internal_faces, nP = g->n_points;
        const int32_t* fo = g->face_offsets;
        const int32_t* fp = g->face_points;
        const int32_t* own = g->owner;
        const int32_t* nei = g->neighbour;
        // ---- ghost layers: layer k = cells sharing a point with a cell of layer < k
        std::vector<int32_t> layer(nC, -1);
        int nOwned = 0;
        for (int c = 0; c < nC; ++c)
            if (cell_rank[c] == rank) { layer[c] = 0; nOwned++; }
        if (nOwned == 0) throw std::invalid_argument("svof_decompose: this rank owns no cells");
        std::vector<unsigned char> mark(nP, 0);
        for (int k = 1; k <= layers; ++k) {
            for (int f = 0; f < nF; ++f) {
                const bool in = layer[own[f]] >= 0 || (f < nIF && layer[nei[f]] >= 0);
                if (in)
                    for (int q = fo[f]; q < fo[f + 1]; ++q) mark[fp[q]] = 1;
            }
            bool grew = false;
            for (int f = 0; f < nF; ++f) {
                const int o = own[f], n = (f < nIF) ? nei[f] : -1;
                if (layer[o] >= 0 && (n < 0 || layer[n] >= 0)) continue;
                bool touch = false;
                for (int q = fo[f]; q < fo[f + 1] && !touch; ++q) touch = mark[fp[q]] != 0;
                if (!touch) continue;
                if (layer[o] < 0) { layer[o] = -2; grew = true; }        // -2: joins layer k (kept apart until the pass is over)
                if (n >= 0 && layer[n] < 0) { layer[n] = -2; grew = true; }
            }
            for (int c = 0; c < nC; ++c)
                if (layer[c] == -2) layer[c] = k;
            if (!grew) break;
        }
        // ---- local cells: ascending global label
        svof_submesh* s = new svof_submesh;
        std::vector<int32_t> cellLocal(nC, -1);
        for (int c = 0; c < nC; ++c)
            if (layer[c] >= 0) {
                cellLocal[c] = (int32_t)s->cellGlobal.size();
                s->cellGlobal.push_back(c);
                s->cellOwnerRank.push_back(cell_rank[c]);
                s->cellLayer.push_back(layer[c]);
                if (layer[c] == 0) s->ownedLocal.push_back(cellLocal[c]);
            }
        s->nCells = (int32_t)s->cellGlobal.size();
        s->nOwned = nOwned;
        // ---- local faces: internal (both sides kept, ascending global label), the physical patches in order, then
        //      the cut faces (one side kept) as a last zeroGradient patch, the kept cell as owner
        std::vector<int32_t> faceList;
        std::vector<unsigned char> flip;
        for (int f = 0; f < nIF; ++f)
            if (layer[own[f]] >= 0 && layer[nei[f]] >= 0) { faceList.push_back(f); flip.push_back(0); }
        s->nInternal = (int32_t)faceList.size();
        s->patches.resize((size_t)g->n_patches + 1);
        for (int pi = 0; pi < g->n_patches; ++pi) {
            svof_patch p = g->patches[pi];
            const int start = (int)faceList.size();
            for (int k = 0; k < g->patches[pi].size; ++k) {
                const int f = g->patches[pi].start + k;
                if (layer[own[f]] >= 0) { faceList.push_back(f); flip.push_back(0); }
            }
            p.start = start;
            p.size = (int)faceList.size() - start;
            s->patches[pi] = p;
        }
        {
            svof_patch p;
            memset(&p, 0, sizeof(p));
            p.start = (int)faceList.size();
            p.kind = SVOF_PATCH_GENERIC;
            p.nbr_rank = -1;
            p.alpha_bc = SVOF_BC_ZERO_GRADIENT;
            for (int f = 0; f < nIF; ++f) {
                const bool a = layer[own[f]] >= 0, b = layer[nei[f]] >= 0;
                if (a != b) { faceList.push_back(f); flip.push_back(b ? 1 : 0); }
            }
            p.size = (int)faceList.size() - p.start;
            s->patches[g->n_patches] = p;
        }
        // ---- local points: ascending global label
        std::fill(mark.begin(), mark.end(), 0);
        for (int f : faceList)
            for (int q = fo[f]; q < fo[f + 1]; ++q) mark[fp[q]] = 1;
        std::vector<int32_t> pointLocal(nP, -1);
        for (int p = 0; p < nP; ++p)
            if (mark[p]) {
                pointLocal[p] = (int32_t)s->pointGlobal.size();
                s->pointGlobal.push_back(p);
                s->points.push_back(g->points[3 * (size_t)p]);
                s->points.push_back(g->points[3 * (size_t)p + 1]);
                s->points.push_back(g->points[3 * (size_t)p + 2]);
            }
        // ---- connectivity
        const size_t nLF = faceList.size();
        s->faceOff.resize(nLF + 1);
        s->owner.resize(nLF);
        s->neighbour.resize((size_t)s->nInternal);
        s->faceGlobal.assign(faceList.begin(), faceList.end());
        s->faceOwnerRank.resize(nLF);
        s->faceFlip.resize(nLF);
        for (size_t i = 0; i < nLF; ++i) {
            s->faceOwnerRank[i] = cell_rank[own[faceList[i]]];
            s->faceFlip[i] = flip[i];
        }
        s->faceOff[0] = 0;
        for (size_t i = 0; i < nLF; ++i) {
            const int f = faceList[i];
            const int a = fo[f], b = fo[f + 1];
            if (!flip[i]) {
                for (int q = a; q < b; ++q) s->facePts.push_back(pointLocal[fp[q]]);
                s->owner[i] = cellLocal[own[f]];
            } else {  // face::reverseFace: keep the first vertex, reverse the rest -- the kept (neighbour) cell becomes the owner
                s->facePts.push_back(pointLocal[fp[a]]);
                for (int q = b - 1; q > a; --q) s->facePts.push_back(pointLocal[fp[q]]);
                s->owner[i] = cellLocal[nei[f]];
            }
            s->faceOff[i + 1] = (int32_t)s->facePts.size();
            if ((int)i < s->nInternal) s->neighbour[i] = cellLocal[nei[f]];
        }
        *out = s;
        return SVOF_OK;
    } catch (const std::invalid_argument& e) {
        g_decompError = e.what();
        return SVOF_ERR_INVALID_ARG;
    } catch (const std::exception& e) {
        g_decompError = e.what();
        return SVOF_ERR_BAD_MESH;
    }
}

int svof_submesh_mesh(const svof_submesh* s, svof_mesh* m)
{
    if (!s || !m) return SVOF_ERR_INVALID_ARG;
    memset(m, 0, sizeof(*m));
    m->n_points = (int32_t)s->pointGlobal.size();
    m->n_faces = (int32_t)s->owner.size();
    m->n_internal_faces = s->nInternal;
    m->n_cells = s->nCells;
    m->n_patches = (int32_t)s->patches.size();
    m->points = s->points.data();
    m->face_offsets = s->faceOff.data();
    m->face_points = s->facePts.data();
    m->owner = s->owner.data();
    m->neighbour = s->neighbour.data();
    m->patches = s->patches.data();
    return SVOF_OK;
}

int svof_submesh_maps(const svof_submesh* s, int32_t* n_owned, const int32_t** cell_global, const int32_t** cell_owner_rank,
                      const int32_t** cell_layer, const int32_t** owned_local, const int32_t** face_global, const int32_t** point_global)
{
    if (!s) return SVOF_ERR_INVALID_ARG;
    if (n_owned) *n_owned = s->nOwned;
    if (cell_global) *cell_global = s->cellGlobal.data();
    if (cell_owner_rank) *cell_owner_rank = s->cellOwnerRank.data();
    if (cell_layer) *cell_layer = s->cellLayer.data();
    if (owned_local) *owned_local = s->ownedLocal.data();
    if (face_global) *face_global = s->faceGlobal.data();
    if (point_global) *point_global = s->pointGlobal.data();
    return SVOF_OK;
}

int svof_submesh_face_maps(const svof_submesh* s, const int32_t** face_owner_rank, const int32_t** face_flip)
{
    if (!s) return SVOF_ERR_INVALID_ARG;
    if (face_owner_rank) *face_owner_rank = s->faceOwnerRank.data();
    if (face_flip) *face_flip = s->faceFlip.data();
    return SVOF_OK;
}

int svof_submesh_free(svof_submesh* s)
{
    delete s;
    return SVOF_OK;
}

}  // extern "C"
