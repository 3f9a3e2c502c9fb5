// svof_b200.cu -- host side of the B200-native SimPLIC step: the handle, the one-off mesh
// precompute, the per-step launch sequence and the C ABI of include/svof.h.
//
// One handle <-> one rank <-> one GPU <-> one CUDA stream.  All numerics run in the kernels of
// svof_kernels.cuh; this file only builds CSR tables, owns device memory and orders launches.
// There is no CPU fallback: every entry point either runs on the device or returns an error.
//
// Build (see geometricvofext_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xcompiler -fPIC -shared
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>   // types only: the library is dlopen()ed at run time (no link-time dependency, loads on CPU-only boxes)

#include "../../include/svof.h"
#include "svof_geom_kernels.cuh"

using namespace svof;

namespace {

thread_local std::string g_createError;

struct Unsupported : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    int kind = -1;  // 0 reconstruct, 1 advect
    bool pending = false;
};

// A few persistent host threads for the gather/scatter loops of svof_step_host (random accesses into field-sized host
// arrays): created once per handle on first use, instead of spawning std::threads in every call.
class HostPool {
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cvWork_, cvDone_;
    std::function<void(int)> fn_;
    std::atomic<unsigned long long> gen_{0};
    int nTasks_ = 0, next_ = 0, pending_ = 0;
    bool stop_ = false;
    // take and run tasks of the current batch until none is left (lock held on entry and on return)
    void drain(std::unique_lock<std::mutex>& lk)
    {
        while (next_ < nTasks_) {
            const int t = next_++;
            lk.unlock();
            fn_(t);
            lk.lock();
            if (--pending_ == 0) cvDone_.notify_all();
        }
    }
    void loop()
    {
        unsigned long long seen = 0;
        for (;;) {
            // the phases of one svof_step_host follow each other within microseconds: spin briefly before sleeping, a
            // condition-variable wake-up costs more than the work of a phase
            for (int i = 0; i < 40000 && gen_.load(std::memory_order_acquire) == seen; ++i) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            std::unique_lock<std::mutex> lk(m_);
            cvWork_.wait(lk, [&] { return stop_ || next_ < nTasks_ || gen_.load(std::memory_order_relaxed) != seen; });
            if (stop_) return;
            seen = gen_.load(std::memory_order_relaxed);
            drain(lk);
        }
    }

public:
    explicit HostPool(int n)
    {
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cvWork_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return (int)th_.size() + 1; }   // the calling thread works too
    // run fn(0..n-1) on the pool's threads and the caller; returns when all are done
    void run(int n, std::function<void(int)> fn)
    {
        if (n <= 0) return;
        std::unique_lock<std::mutex> lk(m_);
        fn_ = std::move(fn);
        nTasks_ = n;
        next_ = 0;
        pending_ = n;
        gen_.fetch_add(1, std::memory_order_release);
        cvWork_.notify_all();
        drain(lk);
        cvDone_.wait(lk, [&] { return pending_ == 0; });
        nTasks_ = 0;
    }
};

}  // namespace

struct svof_handle {
    int device = 0;
    cudaStream_t stream = nullptr;   // main stream: copies, sparse kernels, timing events
    cudaStream_t streamD = nullptr;  // streaming (dense) kernel only: runs concurrently with the sparse chain
    cudaEvent_t evNear = nullptr, evPlic = nullptr, evDense = nullptr, evInputs = nullptr, evCopy = nullptr;
    // zero-copy host step, overlapped form ("zc_overlap" option): U rows pulled on their own stream beside the normals / plane
    // positioning, the streaming kernel forked at the near sets, values that are final after it pushed on its stream
    cudaStream_t streamU = nullptr;
    cudaEvent_t evFront = nullptr, evU = nullptr, evPush = nullptr, evPhiNear = nullptr;
    // bit 0: U rows on their own stream, bit 1: streaming kernel forked at the near sets (default: both); measured and NOT adopted:
    // bit 2: early pushes on the streaming kernel's stream (+0.4 ms), bit 3: phi of all needBounding-cell faces pulled up front
    // (profiles/r5e/r5f/r5g_e2e_ab_matrix.txt)
    int zcOverlap = 3;
    bool zcDenseEarly = false;
    int zcPushCtas = 2;   // resident CTAs per SM of the early push kernels ("zc_push_ctas")
    bool inputsAfterNear = false, freshRecon = false;
    std::map<std::string, double> hostAcc;  // profile: host wall time per phase of svof_step_host (ms)
    std::chrono::steady_clock::time_point hostT;
    int overlap = 1;       // run the streaming kernel on its own low-priority stream, concurrently with the interface kernels
                           // (SVOF_OVERLAP / svof_set_option "overlap"); bitwise the same result, see DESIGN.md section 5
    int advectCount = 0;
    int epochBumps = 0;     // how often the device-side bounding epoch was advanced (tags derive from it)
    int nP = 0, nF = 0, nIF = 0, nC = 0, nBF = 0;
    std::vector<svof_patch> patches;
    svof_params prm;
    StepParams sp;
    MeshDev md;
    int variant = 0;
    int maxCF = 0;
    std::vector<void*> allocs;
    size_t bytes = 0;
    std::string err;

    // fields
    double* alphaBuf[2] = {nullptr, nullptr};
    int cur = 0;
    double* alphaBBuf[2] = {nullptr, nullptr};
    int cb = 0;
    double *phi = nullptr, *alphaPhi = nullptr, *U = nullptr, *Ub = nullptr, *Sp = nullptr, *Su = nullptr;
    double *iN = nullptr, *iD = nullptr, *iC = nullptr, *iS = nullptr, *Un0 = nullptr;
    int *cellSlot = nullptr, *mixedCells = nullptr, *cellStatus = nullptr;
    unsigned int *mixedBits = nullptr, *near1 = nullptr, *near2 = nullptr, *blockSums = nullptr;
    int* near2List = nullptr;
    int2* work = nullptr;
    int capWork = 0, capMixed = 0, capNear = 0, nWords = 0, nScanBlocks = 0;
    double *dVfGeo = nullptr, *dVf = nullptr, *scratchF = nullptr;
    BoundScratch bs;
    int *oobList[2] = {nullptr, nullptr}, *affList = nullptr, *depInit = nullptr, *depLeft = nullptr, *oobIdx = nullptr;
    void* boundRecs = nullptr;
    int capRec = 0;
    unsigned char* oobState = nullptr;
    Ctl* ctl = nullptr;
    Ctl* hctl = nullptr;  // pinned mirror
    PatchDev* dPatches = nullptr;
    // end-to-end transfer reduction (svof_step_host)
    bool sparseIO = true;            // "sparse_io" option
    unsigned int* uBits = nullptr;
    int *uList = nullptr, *hUList = nullptr;
    double *uPacked = nullptr, *hUPacked = nullptr;
    int capU = 0;
    double* alphaPhiPrev = nullptr;  // what the caller's alphaPhi buffer holds
    int *dIdx = nullptr, *hIdx = nullptr;
    double *dVal = nullptr, *hVal = nullptr;
    int capDelta = 0;
    // sparse phi upload (svof_step_host): bitmap over faces of the entries the step can depend on, host gather, device scatter
    // orientationMethod isoRDF (lazily allocated): RDF field, zone bitmap/list, per-mixed-cell work arrays
    double *rdf = nullptr, *rdfB = nullptr, *rdfNormal = nullptr, *rdfRes = nullptr, *rdfAng = nullptr, *rdfSave = nullptr;
    unsigned int* rdfBits = nullptr;
    int* rdfList = nullptr;
    unsigned char* rdfCoarse = nullptr;
    int capRdfList = 0, capRdfMixed = 0, nRdfPrev = 0, rdfIterations = 0;
    std::vector<double> hRdfRes, hRdfAng;
    std::vector<unsigned char> hRdfCoarse;
    int lastNU = 0, lastCntA = 0, lastCntF = 0;   // sizes of the previous step's lists: the speculative read-back sizes
    bool sparsePhi = true;           // "sparse_phi" option (only with sparse_io)
    double sparsePhiTol = 0.0;       // "sparse_phi_exp" option e > 0: cells with |alpha| <= 10^-e count as empty for the phi upload (0: exact)
    // zero-copy host path (svof_step_host with pinned caller buffers): two device face bitmaps (this step's and the previous one's)
    bool zeroCopy = true;            // "zero_copy" option
    unsigned int* zcBits[2] = {nullptr, nullptr};
    int zcCur = 0;
    bool zcPrevValid = false;        // zcBits[zcCur] describes the alphaPhi the caller's buffer holds
    std::map<const void*, void*> pinnedCache;   // caller pointer -> device pointer (nullptr: not device-accessible host memory)
    unsigned int* boundPhiBits = nullptr;   // set by svof_step_host around its doAdvect: see k_bound_deps / k_bound_apply
    const double* boundPhiHost = nullptr;   // zero-copy form: the caller's pinned phi (device address), see loadCellBound
    bool phiBitsReady = false;       // the face bitmap of the CURRENT alpha is already on the host (prefetched by the previous svof_step_host)
    bool phiPartial = false;         // device phi holds stale values on faces between exactly empty cells: not a full field
    int nWordsF = 0, nPhiBlocks = 0;
    size_t capPhiPacked = 0;
    unsigned int *phiBits = nullptr, *hPhiBits2[2] = {nullptr, nullptr};   // host bitmaps double-buffered: this call's and the
    int *phiBlockOff = nullptr, *hPhiBlockOff2[2] = {nullptr, nullptr};    // previous call's / the prefetched next one
    int pb = 0;                      // host buffer holding the bitmap of the current (or last) call
    bool phiPrevBitsValid = false;   // the other host buffer holds the bitmap of the previous sparse call (alphaPhi read-back)
    bool alphaPhiPrevValid = false;  // alphaPhiPrev mirrors the caller's alphaPhi buffer (delta read-back)
    long long phiTotal = 0;          // marked faces of the current bitmap
    std::vector<int> phiCut;         // block boundaries of the equal-share split of the marked faces over the host threads
    double *phiPacked = nullptr, *hPhiPacked = nullptr;
    cudaEvent_t evBits = nullptr;
    HostPool* pool = nullptr;
    const double* hostAlphaSynced = nullptr;     // caller buffers known to hold the previous step's results
    const double* hostAlphaPhiSynced = nullptr;
    long long h2dBytes = 0, d2hBytes = 0;        // bytes actually moved by the last svof_step_host
    DenseStage dstage;
    size_t dstageSmem = 0;
    bool useStaged = false;
    DenseSliced dsliced;             // sliced, transposed rows of the streaming kernel (k_dense_update4); SVOF_DENSE_V4 / "dense_v4"
    bool denseV4 = false;
    bool denseV3 = false;            // k_dense_update3 (batched loads); SVOF_DENSE_V3 / option "dense_v3"
    DenseFast dfast;                 // owner-sorted connectivity of the streaming kernel (k_dense_update2)
    bool boundLanes = true;          // k_bound_run8 (SVOF_BOUND_LANES=0: the sequential walker)
    bool un0Group = false;           // 8 lanes per cut cell for the interface speed (SVOF_UN0=thread: round-1 thread-per-cell kernel)
    int plicCtas = 0;                // "plic_ctas" option: cap on resident CTAs/SM of the plane-positioning kernel (0 = all that fit)
    int denseCtas = 0;               // "dense_ctas" option: cap on resident CTAs/SM of the streaming kernel (0 = no cap)
    int denseSplit = 0;              // "dense_split" option (with dense_ctas): percent of the tiles launched uncapped before plane positioning
    int denseL2 = 0;                 // "dense_l2" option: 1 = L2 evict-first hints in the streaming kernel
    int denseThreads = 256;          // "dense_threads" option: threads per CTA of the capped streaming kernel (128 or 256)
    int forkAt = 1;                  // "fork" option (overlap != 0): 1 = streaming kernel may start after the near sets, 2 = after plane positioning
    int* bPatch = nullptr;
    double* partial = nullptr;
    double* hpartial = nullptr;

    // decomposed runs: NCCL communicator of this handle and the ghost-refresh plan (svof_halo_setup)
    int rank = 0, world = 1;
    ncclComm_t nccl = nullptr;
    struct Plan {   // who sends which local entries to whom, packed per peer
        std::vector<int> peers, sendOff, recvOff;
        int nSend = 0, nRecv = 0;
        int *sendIdx = nullptr, *recvIdx = nullptr;
        double *sendBuf = nullptr, *recvBuf = nullptr, *recvSign = nullptr;   // recvSign: -1 where the local face is flipped
    };
    struct Halo {
        bool active = false, facesReady = false;
        Plan cells, faces;
        int nOwned = 0;
        int* ownedIdx = nullptr;
    } halo;
    bool haveAlpha = false, havePhi = false, haveU = false, bitsValid = false, advected = false;
    bool anyInletOutlet = false;
    unsigned int* calcBitsStore = nullptr;
    unsigned int* calcBits = nullptr;   // overset: one bit per CALCULATED cell (svof_set_cell_types), nullptr = no filter
    bool interfaceDense = false;   // interfaceN/D were set for all cells (svof_set_interface): clear them densely once
    double mapTime = 0;            // alphaMappingTime (reconstruction.C:782)
    bool uPartial = false;   // svof_step_host (sparse_io) uploaded only the rows of U near the interface: not a full field
    double lastDt = 0.0;
    long long launches = 0;
    // svof_step_device: one captured CUDA graph per (alpha buffer parity, patch-value buffer parity, mixed bitmap valid), valid for one dt
    struct StepGraph { cudaGraphExec_t exec = nullptr; double dt = 0; long long nLaunches = 0; int sched = -1; };
    StepGraph graphs[16];   // [schedule slot of the run-time selection][buffer parities, bitmap valid]
    bool capturing = false;
    // run-time schedule selection of svof_step_device ("sched_auto"): the default schedule (slot 0: streaming kernel forked at
    // the near sets, uncapped) against slot 1 (forked after plane positioning, 4 resident CTAs per SM), 6 warm + 12 timed steps each -- which one wins
    // depends on how the interface chain compares with the streaming pass (profiles/r4a, r4b, r4g: 1.10 vs 1.17 ms at 256^3
    // LeVeque; the filled dam-break box prefers the uncapped kernel).  Both give bitwise the same results.
    int tuneMode = 1, tunePhase = 0, tuneSlot = 0;
    bool schedUser = false, retune = false;
    long long tuneSteps = 0;
    int fork0 = 1, dense0 = 0;          // slot 0 as configured
    double tuneMs[2] = {0, 0};
    cudaEvent_t evT[4] = {nullptr, nullptr, nullptr, nullptr};
    double reconTime = 0, advTime = 0, lastReconMs = 0, lastAdvMs = 0;
    double flatMin = 1, flatMax = 1, flatAvg = 1;
    std::vector<EventPair> events;
    cudaEvent_t marks[8] = {nullptr};
    double denseMs = 0;
    long long denseLaunches = 0;
    size_t evNext = 0;
    int sms = 148;
    // SVOF_PROFILE=1: CUDA events around every launch, table printed by svof_destroy
    bool prof = false;
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    std::vector<ProfRec> profRecs;
    std::map<std::string, std::pair<double, long>> profAcc;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw std::runtime_error(buf_);                                                        \
        }                                                                                          \
    } while (0)

template <class T>
T* dalloc(svof_handle* h, size_t n, bool zero = true)
{
    void* p = nullptr;
    const size_t b = std::max<size_t>(n, 1) * sizeof(T);
    CK(cudaMalloc(&p, b));
    if (zero) CK(cudaMemsetAsync(p, 0, b, h->stream));
    h->allocs.push_back(p);
    h->bytes += b;
    return (T*)p;
}
template <class T>
T* dupload(svof_handle* h, const T* src, size_t n)
{
    T* p = dalloc<T>(h, n, false);
    if (n) CK(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return p;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int sparseGrid(svof_handle* h, int threads) { return h->sms * (1024 / threads > 0 ? 1024 / threads : 1); }

void profBegin(svof_handle* h, const char* name, cudaStream_t st);
void profEnd(svof_handle* h, cudaStream_t st);
#define LAUNCH(h, kern, grid, block, ...)                        \
    do {                                                         \
        if ((h)->prof) profBegin(h, #kern, (h)->stream);         \
        kern<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);  \
        if ((h)->prof) profEnd(h, (h)->stream);                  \
        (h)->launches++;                                         \
    } while (0)

// dispatch a capacity-variant launcher (explicitly instantiated in svof_inst.cu)
#define GEO(h, fn, ...)                                               \
    do {                                                              \
        if ((h)->prof) profBegin(h, #fn, (h)->stream);                \
        switch ((h)->variant) {                                       \
            case 0: GeoLaunch<CapsHex>::fn(__VA_ARGS__); break;       \
            case 1: GeoLaunch<CapsSmall>::fn(__VA_ARGS__); break;     \
            case 2: GeoLaunch<CapsPoly>::fn(__VA_ARGS__); break;      \
            case 4: GeoLaunch<CapsHexSplit>::fn(__VA_ARGS__); break;  \
            default: GeoLaunch<CapsSplit>::fn(__VA_ARGS__); break;    \
        }                                                             \
        if ((h)->prof) profEnd(h, (h)->stream);                       \
        (h)->launches++;                                              \
    } while (0)

void profFlush(svof_handle* h)
{
    for (auto& r : h->profRecs) {
        cudaEventSynchronize(r.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& acc = h->profAcc[r.name];
        acc.first += ms;
        acc.second++;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->profRecs.clear();
}
void profBegin(svof_handle* h, const char* name, cudaStream_t st)
{
    if (h->profRecs.size() > 4000) profFlush(h);
    svof_handle::ProfRec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    h->profRecs.push_back(r);
}
void profEnd(svof_handle* h, cudaStream_t st) { cudaEventRecord(h->profRecs.back().b, st); }
void fetchCtl(svof_handle* h);
void profPrint(svof_handle* h)
{
    profFlush(h);
#ifdef SV_BOUND_STATS
    {
        unsigned long long d[8];
        cudaMemcpyFromSymbol(d, g_dbg, sizeof(d));
        fprintf(stderr, "[svof bound stats] cells %llu inner-iterations %llu (%.2f/cell) cycles: load %.0f compute %.0f release %.0f per cell; max chain per thread %llu\n",
                d[0], d[1], (double)d[1] / std::max(1ull, d[0]), (double)d[2] / std::max(1ull, d[0]), (double)d[3] / std::max(1ull, d[0]),
                (double)d[5] / std::max(1ull, d[0]), d[4]);
    }
#endif
    fetchCtl(h);
    {
        const Ctl& c = *h->hctl;
        fprintf(stderr, "[svof profile] last step: nMixed %d nNear2 %d nWork %d | sweep: nAff %d %d %d nPend %d %d %d nearOob %d %d %d %d nOob(left) %d %d\n",
                c.nMixed, c.nNear2, c.nWork, c.nAff[0], c.nAff[1], c.nAff[2], c.nPend[0], c.nPend[1], c.nPend[2], c.nearOob[0],
                c.nearOob[1], c.nearOob[2], c.nearOob[3], c.nOob[0], c.nOob[1]);
    }
    if (!h->hostAcc.empty()) {
        fprintf(stderr, "[svof profile] svof_step_host: host wall time per phase (ms total)\n");
        for (auto& kv : h->hostAcc) fprintf(stderr, "  %-40s %10.3f ms\n", kv.first.c_str(), kv.second);
    }
    double tot = 0;
    for (auto& kv : h->profAcc) tot += kv.second.first;
    fprintf(stderr, "[svof profile] in-situ event timing per launch site (ms total / launches / us each)\n");
    for (auto& kv : h->profAcc)
        fprintf(stderr, "  %-24s %10.3f ms %8ld  %9.2f us  %5.1f%%\n", kv.first.c_str(), kv.second.first, kv.second.second,
                1e3 * kv.second.first / std::max(1L, kv.second.second), 100.0 * kv.second.first / std::max(tot, 1e-30));
}

// Face / cell geometry, face flatness (reconstruction::updateFaceFlatness, reconstruction.C:408-473), tet base points and
// polyMesh::geometricD from the points on the device: at create and again whenever the points move (svof_update_points).
void computeGeometry(svof_handle* h, const double* hCf, const double* hSf, const double* hC, const double* hV)
{
    const MeshDev& d = h->md;
    const int nF = h->nF, nC = h->nC;
    double *Cf = const_cast<double*>(d.Cf), *Sf = const_cast<double*>(d.Sf), *magSf = const_cast<double*>(d.magSf);
    double *C = const_cast<double*>(d.C), *V = const_cast<double*>(d.V), *flat = const_cast<double*>(d.flat);
    unsigned char* tetBase = const_cast<unsigned char*>(d.tetBase);
    const int haveFaceGeom = (hCf && hSf) ? 1 : 0;
    if (haveFaceGeom) {
        CK(cudaMemcpyAsync(Cf, hCf, sizeof(double) * 3 * nF, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(Sf, hSf, sizeof(double) * 3 * nF, cudaMemcpyHostToDevice, h->stream));
    }
    LAUNCH(h, k_face_geom, cdiv(nF, 256), 256, d, Cf, Sf, magSf, haveFaceGeom);
    if (hC && hV) {
        CK(cudaMemcpyAsync(C, hC, sizeof(double) * 3 * nC, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(V, hV, sizeof(double) * nC, cudaMemcpyHostToDevice, h->stream));
    } else {
        LAUNCH(h, k_cell_geom, cdiv(nC, 256), 256, d, C, V);
    }
    LAUNCH(h, k_flatness_tetbase, cdiv(nF, 256), 256, d, flat, tetBase);
    CK(cudaStreamSynchronize(h->stream));

    // "SimPLIC::Mesh face flatness: min/max/avg" (reconstruction.C:442-447)
    {
        std::vector<double> hf(nF), hm(nF);
        CK(cudaMemcpy(hf.data(), flat, sizeof(double) * nF, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hm.data(), magSf, sizeof(double) * nF, cudaMemcpyDeviceToHost));
        double mn = SV_VGREAT, mx = -SV_VGREAT, sfa = 0, sa = 0;
        for (int f = 0; f < nF; ++f) {
            mn = std::min(mn, hf[f]); mx = std::max(mx, hf[f]);
            sfa += hf[f] * hm[f]; sa += hm[f];
        }
        h->flatMin = mn; h->flatMax = mx; h->flatAvg = sfa / sa;
    }

    // polyMesh::geometricD(): directions normal to empty patches are excluded (OF, recalled)
    h->sp.geomD[0] = h->sp.geomD[1] = h->sp.geomD[2] = 1;
    {
        bool anyEmpty = false;
        for (const svof_patch& p : h->patches) anyEmpty |= (p.kind == SVOF_PATCH_EMPTY && p.size > 0);
        if (anyEmpty) {
            std::vector<double> hs((size_t)3 * nF);
            CK(cudaMemcpy(hs.data(), Sf, sizeof(double) * 3 * nF, cudaMemcpyDeviceToHost));
            double e[3] = {0, 0, 0};
            for (const svof_patch& p : h->patches) {
                if (p.kind != SVOF_PATCH_EMPTY) continue;
                for (int k = 0; k < p.size; ++k) {
                    const double* s = &hs[(size_t)3 * (p.start + k)];
                    const double ms = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
                    if (ms > 0) for (int q = 0; q < 3; ++q) e[q] += std::fabs(s[q] / ms);
                }
            }
            const double me = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
            if (me > 0) for (int q = 0; q < 3; ++q) h->sp.geomD[q] = (e[q] / me > 1e-6) ? -1 : 1;
        }
    }
}

void buildMesh(svof_handle* h, const svof_mesh& m)
{
    if (!m.points || !m.face_offsets || !m.face_points || !m.owner || (m.n_internal_faces > 0 && !m.neighbour) ||
        (m.n_patches > 0 && !m.patches))
        throw std::invalid_argument("svof_mesh: null connectivity pointer");
    const int nP = m.n_points, nF = m.n_faces, nIF = m.n_internal_faces, nC = m.n_cells, nBF = nF - nIF;
    if (nP <= 0 || nF <= 0 || nC <= 0 || nIF < 0 || nIF > nF) throw std::invalid_argument("svof_mesh: bad sizes");
    h->nP = nP; h->nF = nF; h->nIF = nIF; h->nC = nC; h->nBF = nBF;
    h->patches.assign(m.patches, m.patches + m.n_patches);
    const int* fo = m.face_offsets;
    const int* fp = m.face_points;
    const int* own = m.owner;
    const int* nei = m.neighbour;
    int maxFV = 0;
    for (int f = 0; f < nF; ++f) {
        const int nv = fo[f + 1] - fo[f];
        if (nv < 3) throw std::invalid_argument("svof_mesh: face with < 3 points");
        if (own[f] < 0 || own[f] >= nC) throw std::invalid_argument("svof_mesh: owner out of range");
        if (f < nIF && (nei[f] < 0 || nei[f] >= nC)) throw std::invalid_argument("svof_mesh: neighbour out of range");
        maxFV = std::max(maxFV, nv);
    }
    if (maxFV > 255) throw std::invalid_argument("svof_mesh: face with > 255 points");
    const long long nFP = fo[nF];
    for (long long i = 0; i < nFP; ++i)
        if (fp[i] < 0 || fp[i] >= nP) throw std::invalid_argument("svof_mesh: point label out of range");

    // patch table must tile the boundary faces in order
    std::vector<unsigned char> bKind(std::max(nBF, 1), 0);
    std::vector<int> bPatch(std::max(nBF, 1), 0);
    std::vector<PatchDev> pd(std::max<size_t>(h->patches.size(), 1));
    int expect = nIF;
    for (size_t pi = 0; pi < h->patches.size(); ++pi) {
        const svof_patch& p = h->patches[pi];
        if (p.start != expect || p.size < 0 || p.start + p.size > nF)
            throw std::invalid_argument("svof_mesh: patches must tile the boundary faces in order");
        if (p.kind < 0 || p.kind > 2) throw std::invalid_argument("svof_mesh: unknown patch kind");
        // advection.C:311-393 (syncProcPatches) is replaced by ghost layers: a decomposed run hands every rank its cells plus
        // halo layers (svof_decompose) and the library refreshes them (svof_halo_*); raw processor patches would advect
        // alpha = 0 through the cut, so they are refused instead of silently accepted
        if (p.kind == SVOF_PATCH_PROCESSOR && p.size > 0)
            throw Unsupported("svof_mesh: non-empty processor patch: decomposed runs go through svof_decompose + svof_halo_setup "
                              "(ghost layers), not through processor patches");
        for (int k = 0; k < p.size; ++k) {
            bKind[p.start - nIF + k] = (unsigned char)p.kind;
            bPatch[p.start - nIF + k] = (int)pi;
        }
        pd[pi].start = p.start; pd[pi].size = p.size; pd[pi].kind = p.kind; pd[pi].bc = p.alpha_bc; pd[pi].value = p.alpha_value;
        if (p.kind == SVOF_PATCH_GENERIC && p.alpha_bc == SVOF_BC_INLET_OUTLET && p.size > 0) h->anyInletOutlet = true;
        expect += p.size;
    }
    if (expect != nF) throw std::invalid_argument("svof_mesh: patches do not cover all boundary faces");

    // cells(): owned faces ascending, then neighbour-side faces ascending (primitiveMesh::calcCells)
    std::vector<int> cellOff(nC + 1, 0);
    for (int f = 0; f < nF; ++f) cellOff[own[f] + 1]++;
    for (int f = 0; f < nIF; ++f) cellOff[nei[f] + 1]++;
    for (int c = 0; c < nC; ++c) cellOff[c + 1] += cellOff[c];
    const int nCF = cellOff[nC];
    std::vector<int> cellFaces(nCF);
    std::vector<int> nOwned(nC, 0);
    {
        std::vector<int> fill(cellOff.begin(), cellOff.end() - 1);
        for (int f = 0; f < nF; ++f) { cellFaces[fill[own[f]]++] = f; nOwned[own[f]]++; }
        for (int f = 0; f < nIF; ++f) cellFaces[fill[nei[f]]++] = f;
    }
    // ascending-face rows for the streaming kernel: merge of the two sorted runs
    std::vector<int2> cellAsc(nCF);
    int maxCF = 0;
    for (int c = 0; c < nC; ++c) {
        const int b = cellOff[c], e = cellOff[c + 1], mid = b + nOwned[c];
        maxCF = std::max(maxCF, e - b);
        int i = b, j = mid, o = b;
        while (i < mid || j < e) {
            const bool takeOwned = (j >= e) || (i < mid && cellFaces[i] < cellFaces[j]);
            if (takeOwned) {
                const int f = cellFaces[i++];
                cellAsc[o++] = make_int2(f, f < nIF ? nei[f] : (-1 - (f - nIF)));
            } else {
                const int f = cellFaces[j++];
                cellAsc[o++] = make_int2(f | (int)0x80000000, own[f]);
            }
        }
    }
    // staging plan of the streaming kernel: per 256-cell CTA, the contiguous range of internal faces its cells own
    // (internal faces are ordered by owner in any OpenFOAM mesh) and the shared-memory capacities
    std::vector<int> ctaFace;
    int maxRowSlab = 0, maxPhiSlab = 0;
    {
        bool ownerSorted = true;
        for (int f = 1; f < nIF; ++f) ownerSorted &= (own[f - 1] <= own[f]);
        const int nCta = (nC + 255) / 256;
        if (ownerSorted) {
            ctaFace.resize(nCta + 1);
            int f = 0;
            for (int b = 0; b <= nCta; ++b) {
                const int cFirst = std::min(b * 256, nC);
                while (f < nIF && own[f] < cFirst) ++f;
                ctaFace[b] = f;
            }
        }
        for (int b = 0; b < nCta; ++b) {
            const int cA = b * 256, cB = std::min(nC, cA + 256);
            maxRowSlab = std::max(maxRowSlab, cellOff[cB] - (cellOff[cA] & ~1));
            if (ownerSorted) maxPhiSlab = std::max(maxPhiSlab, ctaFace[b + 1] - (ctaFace[b] & ~1));
        }
    }
    // owner-sorted connectivity of the streaming kernel (k_dense_update2): count byte per cell, {first owned internal
    // face, first neighbour-side row} per 32 cells, {face, owner} rows of the neighbour-side faces
    std::vector<unsigned char> dfCnt;
    std::vector<int2> dfWarpBase, dfLowRows;
    std::vector<unsigned int> dfSlow;
    // measured at 256^3 (profiles/r2c_*): 568 us against 432 us for the int2-row kernel (the batched, unconditional neighbour
    // loads spill at 32 registers and double the alpha gathers) -> opt-in (SVOF_DENSE_V2=1) until that is fixed
    bool dfOk = getenv("SVOF_DENSE_V2") != nullptr && atoi(getenv("SVOF_DENSE_V2")) > 0;
    {
        for (int f = 1; f < nIF && dfOk; ++f) dfOk = (own[f - 1] <= own[f]);
        std::vector<int> nOwnI(nC, 0), nLow(nC, 0);
        for (int f = 0; f < nIF; ++f) { nOwnI[own[f]]++; nLow[nei[f]]++; }
        for (int c = 0; c < nC && dfOk; ++c) dfOk = (nOwnI[c] <= 15 && nLow[c] <= 15);
        if (dfOk) {
            const int nW = (nC + 31) / 32;
            dfCnt.resize(nC);
            dfWarpBase.resize(nW);
            dfSlow.assign(nW, 0u);
            dfLowRows.resize(std::max(nIF, 1));
            std::vector<int> lowOff(nC + 1, 0);
            int fo0 = 0;
            for (int c = 0; c < nC; ++c) {
                if ((c & 31) == 0) dfWarpBase[c >> 5] = make_int2(fo0, lowOff[c]);
                dfCnt[c] = (unsigned char)(nOwnI[c] | (nLow[c] << 4));
                fo0 += nOwnI[c];
                lowOff[c + 1] = lowOff[c] + nLow[c];
            }
            std::vector<int> fill(lowOff.begin(), lowOff.end() - 1);
            for (int f = 0; f < nIF; ++f) dfLowRows[fill[nei[f]]++] = make_int2(f, own[f]);  // ascending face within each cell
            for (int f = nIF; f < nF; ++f) dfSlow[own[f] >> 5] |= 1u << (own[f] & 31);        // cells with boundary faces: generic rows
        }
    }
    // cellPoints ascending; pointCells ascending.  The per-cell sort/unique is the slowest host loop of the build
    // (21-27 s of set-up at 256^3 in round 1): host threads over cell ranges, counts first, then the fill.
    std::vector<int> cellPtOff(nC + 1, 0), cellPts;
    int maxCP = 0;
    {
        const unsigned nT = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        auto uniquePts = [&](int c, std::vector<int>& tmp) {
            tmp.clear();
            for (int k = cellOff[c]; k < cellOff[c + 1]; ++k) {
                const int f = cellFaces[k];
                tmp.insert(tmp.end(), fp + fo[f], fp + fo[f + 1]);
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        };
        auto forRanges = [&](auto&& body) {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nT; ++t) {
                const int c0 = (int)((long long)nC * t / nT), c1 = (int)((long long)nC * (t + 1) / nT);
                th.emplace_back([&, c0, c1, t]() { body(c0, c1, t); });
            }
            for (auto& x : th) x.join();
        };
        std::vector<int> maxPer(nT, 0);
        forRanges([&](int c0, int c1, unsigned t) {
            std::vector<int> tmp;
            for (int c = c0; c < c1; ++c) {
                uniquePts(c, tmp);
                cellPtOff[c + 1] = (int)tmp.size();
                maxPer[t] = std::max(maxPer[t], (int)tmp.size());
            }
        });
        for (unsigned t = 0; t < nT; ++t) maxCP = std::max(maxCP, maxPer[t]);
        for (int c = 0; c < nC; ++c) cellPtOff[c + 1] += cellPtOff[c];
        cellPts.resize((size_t)cellPtOff[nC]);
        forRanges([&](int c0, int c1, unsigned) {
            std::vector<int> tmp;
            for (int c = c0; c < c1; ++c) {
                uniquePts(c, tmp);
                std::copy(tmp.begin(), tmp.end(), cellPts.begin() + cellPtOff[c]);
            }
        });
    }
    std::vector<int> ptCellOff(nP + 1, 0);
    for (int p : cellPts) ptCellOff[p + 1]++;
    for (int p = 0; p < nP; ++p) ptCellOff[p + 1] += ptCellOff[p];
    std::vector<int> ptCells(ptCellOff[nP]);
    {
        std::vector<int> fill(ptCellOff.begin(), ptCellOff.end() - 1);
        for (int c = 0; c < nC; ++c)
            for (int k = cellPtOff[c]; k < cellPtOff[c + 1]; ++k) ptCells[fill[cellPts[k]]++] = c;
    }
    // boundary point -> boundary faces (ascending), patch points
    std::vector<int> ptBFOff(nP + 1, 0);
    std::vector<unsigned char> isPatchPoint(nP, 0);
    for (int bf = 0; bf < nBF; ++bf) {
        const int f = nIF + bf;
        for (int k = fo[f]; k < fo[f + 1]; ++k) {
            ptBFOff[fp[k] + 1]++;
            if (bKind[bf] == SVOF_PATCH_GENERIC) isPatchPoint[fp[k]] = 1;
        }
    }
    for (int p = 0; p < nP; ++p) ptBFOff[p + 1] += ptBFOff[p];
    std::vector<int> ptBFaces(std::max(ptBFOff[nP], 1));
    {
        std::vector<int> fill(ptBFOff.begin(), ptBFOff.end() - 1);
        for (int bf = 0; bf < nBF; ++bf) {
            const int f = nIF + bf;
            for (int k = fo[f]; k < fo[f + 1]; ++k) ptBFaces[fill[fp[k]]++] = bf;
        }
    }

    // capacity variant
    const bool split = h->prm.split_warped_face != 0;
    int maxLocalFaces = maxCF, maxLocalPts = maxCP;
    if (split) {  // worst case: every face triangulated
        maxLocalFaces = 0;
        for (int c = 0; c < nC; ++c) {
            int s = 0;
            for (int k = cellOff[c]; k < cellOff[c + 1]; ++k) s += fo[cellFaces[k] + 1] - fo[cellFaces[k]];
            maxLocalFaces = std::max(maxLocalFaces, s);
        }
        maxLocalPts = maxCP + maxCF;
    }
    auto fits = [&](int fv, int cf, int cp) { return maxFV <= fv && maxLocalFaces <= cf && maxLocalPts <= cp; };
    if (fits(CapsHex::MAXFV, CapsHex::MAXCF, CapsHex::MAXCP)) h->variant = 0;
    else if (fits(CapsSmall::MAXFV, CapsSmall::MAXCF, CapsSmall::MAXCP)) h->variant = 1;
    else if (fits(CapsHexSplit::MAXFV, CapsHexSplit::MAXCF, CapsHexSplit::MAXCP)) h->variant = 4;
    else if (fits(CapsPoly::MAXFV, CapsPoly::MAXCF, CapsPoly::MAXCP)) h->variant = 2;
    else if (fits(CapsSplit::MAXFV, CapsSplit::MAXCF, CapsSplit::MAXCP)) h->variant = 3;
    else {
        char b[256];
        snprintf(b, sizeof(b), "mesh exceeds the compiled polyhedron caps: face verts %d, cell faces %d, cell points %d",
                 maxFV, maxLocalFaces, maxLocalPts);
        throw std::length_error(b);
    }
    if (maxCF > 64) throw std::length_error("cells with more than 64 faces are not supported");
    h->maxCF = maxCF;

    // upload
    MeshDev& d = h->md;
    d.nPoints = nP; d.nFaces = nF; d.nIF = nIF; d.nCells = nC; d.nBF = nBF;
    d.maxFV = maxFV; d.maxLocalFaces = maxLocalFaces; d.maxLocalPts = maxLocalPts;
    d.points = dupload(h, m.points, (size_t)3 * nP);
    d.faceOff = dupload(h, fo, (size_t)nF + 1);
    d.facePts = dupload(h, fp, (size_t)nFP);
    d.owner = dupload(h, own, (size_t)nF);
    d.neighbour = dupload(h, nei, (size_t)nIF);
    d.cellOff = dupload(h, cellOff.data(), cellOff.size());
    d.cellFaces = dupload(h, cellFaces.data(), cellFaces.size());
    cellAsc.push_back(make_int2(0, -1));  // slack: the staged streaming kernel copies rows in 16-byte requests
    cellAsc.push_back(make_int2(0, -1));
    d.cellAsc = dupload(h, cellAsc.data(), cellAsc.size());
    d.cellPtOff = dupload(h, cellPtOff.data(), cellPtOff.size());
    d.cellPts = dupload(h, cellPts.data(), cellPts.size());
    d.ptCellOff = dupload(h, ptCellOff.data(), ptCellOff.size());
    d.ptCells = dupload(h, ptCells.data(), ptCells.size());
    d.ptBFOff = dupload(h, ptBFOff.data(), ptBFOff.size());
    d.ptBFaces = dupload(h, ptBFaces.data(), ptBFaces.size());
    d.bKind = dupload(h, bKind.data(), bKind.size());
    d.isPatchPoint = dupload(h, isPatchPoint.data(), isPatchPoint.size());
    h->dPatches = dupload(h, pd.data(), pd.size());
    h->dstage.ctaFace = ctaFace.empty() ? nullptr : dupload(h, ctaFace.data(), ctaFace.size());
    h->dstage.rowCap = (std::min(maxRowSlab + 2, 4096) + 1) & ~1;   // <= 32 KB of rows per CTA, even
    h->dstage.phiCap = (std::min(maxPhiSlab + 2, 1536) + 1) & ~1;   // <= 12 KB of phi per CTA
    h->dstageSmem = (size_t)(h->dstage.rowCap + 2) * 8 + (size_t)(h->dstage.phiCap + 2) * 8;
    // measured (profiles/r1l_*): 0.431 ms for the CSR kernel vs 0.556 ms staged -> staged is opt-in (SVOF_DENSE_STAGED=1)
    h->useStaged = getenv("SVOF_DENSE_STAGED") && atoi(getenv("SVOF_DENSE_STAGED")) > 0;
    h->denseV3 = getenv("SVOF_DENSE_V3") && atoi(getenv("SVOF_DENSE_V3")) > 0;
    h->denseV4 = getenv("SVOF_DENSE_V4") && atoi(getenv("SVOF_DENSE_V4")) > 0;
    h->dsliced.enabled = 0;
    if (h->denseV4) {   // sliced, transposed rows: per 32 consecutive cells, entry q of all 32 rows contiguous, padded to the widest row
        // measured at 256^3 (profiles/r2w_dense_variants.txt): 504 us against 426 us for the strided int2 rows -- the strided rows
        // act as a prefetch (one DRAM access brings the lines of the next five iterations into L1) -> opt-in, built on request only
        const int nSl = (nC + 31) / 32;
        std::vector<int> sliceOff(nSl + 1, 0);
        for (int sl = 0; sl < nSl; ++sl) {
            int w = 0;
            for (int c = sl * 32; c < std::min(nC, sl * 32 + 32); ++c) w = std::max(w, cellOff[c + 1] - cellOff[c]);
            sliceOff[sl + 1] = sliceOff[sl] + w;
        }
        std::vector<int2> rowsT((size_t)sliceOff[nSl] * 32, make_int2(SV_ROW_PAD, SV_ROW_PAD));
        for (int c = 0; c < nC; ++c) {
            const size_t base = (size_t)sliceOff[c >> 5] * 32 + (c & 31);
            for (int k = cellOff[c]; k < cellOff[c + 1]; ++k) rowsT[base + (size_t)(k - cellOff[c]) * 32] = cellAsc[k];
        }
        h->dsliced.rowsT = dupload(h, rowsT.data(), rowsT.size());
        h->dsliced.sliceOff = dupload(h, sliceOff.data(), sliceOff.size());
        h->dsliced.enabled = 1;
        CK(cudaStreamSynchronize(h->stream));
    }
    CK(cudaFuncSetAttribute(k_dense_update_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->dstageSmem));
    h->bPatch = dupload(h, bPatch.data(), bPatch.size());
    h->dfast.enabled = dfOk ? 1 : 0;
    if (dfOk) {
        h->dfast.cnt = dupload(h, dfCnt.data(), dfCnt.size());
        h->dfast.warpBase = dupload(h, dfWarpBase.data(), dfWarpBase.size());
        h->dfast.lowRows = dupload(h, dfLowRows.data(), dfLowRows.size());
        h->dfast.slowBits = dupload(h, dfSlow.data(), dfSlow.size());
    }

    d.Cf = dalloc<double>(h, (size_t)3 * nF);
    d.Sf = dalloc<double>(h, (size_t)3 * nF);
    d.magSf = dalloc<double>(h, nF);
    d.C = dalloc<double>(h, (size_t)3 * nC);
    d.V = dalloc<double>(h, nC);
    d.flat = dalloc<double>(h, nF);
    d.tetBase = dalloc<unsigned char>(h, nF);
    CK(cudaStreamSynchronize(h->stream));  // host vectors go out of scope
    computeGeometry(h, m.Cf, m.Sf, m.C, m.V);

    // work-list capacities
    h->capMixed = nC;
    h->capNear = nC;
    h->capWork = (int)std::min<long long>(nCF, std::max<long long>(1 << 20, nCF / 3));
    h->nWords = cdiv(nC, 32);
    h->nScanBlocks = cdiv(h->nWords, SV_SCAN_WORDS);
}

void allocFields(svof_handle* h)
{
    const size_t nC = h->nC, nF = h->nF, nBF = std::max(h->nBF, 1);
    h->alphaBuf[0] = dalloc<double>(h, nC);
    h->alphaBuf[1] = dalloc<double>(h, nC);
    h->phi = dalloc<double>(h, nF + 2);  // +2: the staged streaming kernel may read one 16-byte request past nIF
    h->alphaPhi = dalloc<double>(h, nF);
    h->alphaBBuf[0] = dalloc<double>(h, nBF);
    h->alphaBBuf[1] = dalloc<double>(h, nBF);
    h->U = dalloc<double>(h, 3 * nC);
    h->Ub = dalloc<double>(h, 3 * nBF);
    h->Sp = dalloc<double>(h, nC);
    h->Su = dalloc<double>(h, nC);
    h->iN = dalloc<double>(h, 3 * nC);
    h->iD = dalloc<double>(h, nC);
    h->iC = dalloc<double>(h, 3 * nC);
    h->iS = dalloc<double>(h, 3 * nC);
    h->cellSlot = dalloc<int>(h, nC, false);
    CK(cudaMemsetAsync(h->cellSlot, 0xff, nC * sizeof(int), h->stream));
    h->mixedCells = dalloc<int>(h, h->capMixed);
    h->cellStatus = dalloc<int>(h, h->capMixed);
    h->Un0 = dalloc<double>(h, h->capMixed);
    h->mixedBits = dalloc<unsigned int>(h, h->nWords + 1);
    h->near1 = dalloc<unsigned int>(h, h->nWords + 1);
    h->near2 = dalloc<unsigned int>(h, h->nWords + 1);
    h->blockSums = dalloc<unsigned int>(h, h->nScanBlocks + 1);
    h->near2List = dalloc<int>(h, h->capNear);
    h->work = dalloc<int2>(h, h->capWork);
    h->dVfGeo = dalloc<double>(h, nF);
    h->dVf = dalloc<double>(h, nF);
    h->scratchF = dalloc<double>(h, nF);
    h->bs.corr = dalloc<double>(h, nF);
    h->bs.tagV = dalloc<int>(h, nF);
    h->bs.tagR = dalloc<int>(h, nF);
    h->bs.corrBy = dalloc<int>(h, nF);
    h->bs.corrPos = dalloc<int>(h, nF);
    h->bs.affStamp = dalloc<int>(h, nC);
    h->oobList[0] = dalloc<int>(h, h->capNear);
    h->oobList[1] = dalloc<int>(h, h->capNear);
    h->affList = dalloc<int>(h, h->capNear);
    h->depInit = dalloc<int>(h, nC);
    h->depLeft = dalloc<int>(h, nC);
    h->oobIdx = dalloc<int>(h, nC);
    {
        const size_t recBytes = (h->maxCF <= 8) ? sizeof(CellBound<8>) : (h->maxCF <= 16) ? sizeof(CellBound<16>) : sizeof(CellBound<64>);
        h->capRec = (int)std::min<size_t>(nC, std::max<size_t>(65536, nC / 8));   // out-of-bounds cells per sweep; overflow raises SVERR_LIST (returned at the next sync point)
        h->boundRecs = dalloc<unsigned char>(h, recBytes * (size_t)h->capRec, false);
    }
    h->oobState = dalloc<unsigned char>(h, nC);
    h->ctl = dalloc<Ctl>(h, 1);
    h->uBits = dalloc<unsigned int>(h, h->nWords + 1);
    h->capU = (int)std::min<size_t>(nC, std::max<size_t>(1 << 20, nC / 8));
    h->uList = dalloc<int>(h, h->capU);
    h->uPacked = dalloc<double>(h, 3 * (size_t)h->capU);
    h->alphaPhiPrev = dalloc<double>(h, nF);
    h->capDelta = (int)std::min<size_t>(nF, std::max<size_t>(1 << 20, nF / 6));
    h->dIdx = dalloc<int>(h, h->capDelta);
    h->dVal = dalloc<double>(h, h->capDelta);
    CK(cudaMallocHost((void**)&h->hUList, sizeof(int) * h->capU));
    CK(cudaMallocHost((void**)&h->hUPacked, sizeof(double) * 3 * (size_t)h->capU));
    CK(cudaMallocHost((void**)&h->hIdx, sizeof(int) * h->capDelta));
    CK(cudaMallocHost((void**)&h->hVal, sizeof(double) * h->capDelta));
    h->partial = dalloc<double>(h, 1024);
    CK(cudaMallocHost((void**)&h->hctl, sizeof(Ctl)));
    CK(cudaMallocHost((void**)&h->hpartial, 1024 * sizeof(double)));
    memset(h->hctl, 0, sizeof(Ctl));
    if (h->evNear) return;   // svof_update_mesh: the events of the handle are kept
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&h->marks[i]));
    CK(cudaEventCreateWithFlags(&h->evNear, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evPlic, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evDense, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evInputs, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evCopy, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evFront, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evU, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evPush, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evPhiNear, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&h->evT[i]));
    h->events.resize(96);
    for (EventPair& e : h->events) {
        CK(cudaEventCreate(&e.a));
        CK(cudaEventCreate(&e.b));
    }
}

void harvestEvents(svof_handle* h, bool block)
{
    for (EventPair& e : h->events) {
        if (!e.pending) continue;
        if (block) CK(cudaEventSynchronize(e.b));
        else if (cudaEventQuery(e.b) != cudaSuccess) { (void)cudaGetLastError(); continue; }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e.a, e.b));
        if (e.kind == 0) { h->reconTime += ms * 1e-3; h->lastReconMs = ms; }
        else if (e.kind == 1) { h->advTime += ms * 1e-3; h->lastAdvMs = ms; }
        else { h->denseMs += ms; h->denseLaunches++; }
        e.pending = false;
    }
}
EventPair& beginTimedOn(svof_handle* h, int kind, cudaStream_t st)
{
    EventPair& e = h->events[h->evNext++ % h->events.size()];
    if (e.pending) {
        CK(cudaEventSynchronize(e.b));
        harvestEvents(h, false);
    }
    e.kind = kind;
    CK(cudaEventRecord(e.a, st));
    return e;
}
void endTimedOn(svof_handle* h, EventPair& e, cudaStream_t st)
{
    CK(cudaEventRecord(e.b, st));
    e.pending = true;
}
EventPair& beginTimed(svof_handle* h, int kind)
{
    EventPair& e = h->events[h->evNext++ % h->events.size()];
    if (e.pending) {
        CK(cudaEventSynchronize(e.b));
        harvestEvents(h, false);
    }
    e.kind = kind;
    CK(cudaEventRecord(e.a, h->stream));
    return e;
}
void endTimed(svof_handle* h, EventPair& e)
{
    CK(cudaEventRecord(e.b, h->stream));
    e.pending = true;
}

__global__ void k_ctl_reset_advect(Ctl* ctl)
{
    ctl->nWork = 0;
    ctl->epoch++;
    ctl->nOob[0] = ctl->nOob[1] = 0;
    for (int s = 0; s <= SV_MAX_SWEEPS; ++s) ctl->nPend[s] = ctl->nAff[s] = ctl->nearOob[s] = 0;
    ctl->minNear0 = ctl->minNearF = ~0ull;
    ctl->maxNear0 = ctl->maxNearF = 0ull;
}
__global__ void k_ctl_reset_dense(Ctl* ctl)
{
    ctl->minDense = ~0ull;
    ctl->maxDense = 0ull;
}


// ---- NCCL, bound at run time ------------------------------------------------------------------------
// dlopen instead of a link-time dependency: the library must load on CPU-only boxes (symbol checks, svof_decompose),
// and a host process that already carries an NCCL (torch.distributed, an MPI build) must not end up with two copies.
struct NcclApi {
    bool tried = false, ok = false;
    std::string why;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& ncclApi()
{
    static NcclApi a;
    if (a.tried) return a;
    a.tried = true;
    void* lib = nullptr;
    const char* names[] = {getenv("SVOF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {   // an NCCL already in the process (torch's) wins over loading another one
        if (n && !lib) lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    }
    for (const char* n : names) {
        if (n && !lib) lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!lib) {
        a.why = std::string("libnccl.so.2 not loadable: ") + (dlerror() ? dlerror() : "?");
        return a;
    }
#define SV_NCCL_SYM(field, name)                                   \
    a.field = (decltype(a.field))dlsym(lib, name);                 \
    if (!a.field) { a.why = std::string("missing symbol ") + name; return a; }
    SV_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    SV_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    SV_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    SV_NCCL_SYM(GroupStart, "ncclGroupStart")
    SV_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    SV_NCCL_SYM(Send, "ncclSend")
    SV_NCCL_SYM(Recv, "ncclRecv")
    SV_NCCL_SYM(AllGather, "ncclAllGather")
    SV_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SV_NCCL_SYM
    a.ok = true;
    return a;
}
struct CommError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define NK(call)                                                                                  \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess) {                                                                  \
            char buf_[512];                                                                       \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, ncclApi().GetErrorString(r_), __FILE__, __LINE__); \
            throw CommError(buf_);                                                                \
        }                                                                                         \
    } while (0)

__global__ void k_gather_alpha(const int* __restrict__ idx, int n, const double* __restrict__ alpha, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = alpha[idx[i]];
}
__global__ void k_volume_partial_list(const int* __restrict__ idx, int n, const double* alpha, const double* V, double* partial)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += alpha[idx[i]] * V[idx[i]];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ghost alpha <- owning rank: pack, one grouped NCCL send/recv, scatter (keeps the mixed-cell bitmap valid).  Everything
// is enqueued on the handle's stream: no host synchronisation.  This one exchange stands where the reference has the
// zoneDistribute stencil exchange (reconstruction.C:97-107) and 1 + 2 x sweeps calls of syncProcPatches (advection.C:311-393).
__global__ void k_gather_comps(const int* __restrict__ idx, int n, int comps, const double* __restrict__ src, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * comps) return;
    out[i] = src[(size_t)idx[i / comps] * comps + (i % comps)];
}
__global__ void k_scatter_comps(const int* __restrict__ idx, int n, int comps, const double* __restrict__ vals, const double* __restrict__ sign,
                                double* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * comps) return;
    const double v = vals[i];
    dst[(size_t)idx[i / comps] * comps + (i % comps)] = sign ? sign[i / comps] * v : v;
}

void planSendRecv(svof_handle* h, const svof_handle::Plan& P, int comps)
{
    NcclApi& N = ncclApi();
    NK(N.GroupStart());
    for (size_t p = 0; p < P.peers.size(); ++p) {
        const int ns = P.sendOff[p + 1] - P.sendOff[p], nr = P.recvOff[p + 1] - P.recvOff[p];
        if (ns) NK(N.Send(P.sendBuf + (size_t)P.sendOff[p] * comps, (size_t)ns * comps, ncclDouble, P.peers[p], h->nccl, h->stream));
        if (nr) NK(N.Recv(P.recvBuf + (size_t)P.recvOff[p] * comps, (size_t)nr * comps, ncclDouble, P.peers[p], h->nccl, h->stream));
    }
    NK(N.GroupEnd());
}

// ghost alpha <- owning rank: pack, one grouped NCCL send/recv, scatter (keeps the mixed-cell bitmap valid).  Everything
// is enqueued on the handle's stream: no host synchronisation.  This one exchange stands where the reference has the
// zoneDistribute stencil exchange (reconstruction.C:97-107) and 1 + 2 x sweeps calls of syncProcPatches (advection.C:311-393).
void haloExchange(svof_handle* h)
{
    if (!h->halo.active) return;
    const svof_handle::Plan& P = h->halo.cells;
    double* alpha = h->alphaBuf[h->cur];
    if (P.nSend) LAUNCH(h, k_gather_alpha, cdiv(P.nSend, 256), 256, P.sendIdx, P.nSend, alpha, P.sendBuf);
    planSendRecv(h, P, 1);
    if (P.nRecv)
        LAUNCH(h, k_scatter_alpha, cdiv(P.nRecv, 256), 256, P.recvIdx, P.recvBuf, (long long)P.nRecv, h->prm.mixed_cell_tol, alpha,
               h->mixedBits, h->bitsValid ? 1 : 0);
}

// ghost U (cells, 3 components) and phi (faces) <- owning ranks: what a flow solver that only holds its own cells needs
// before the step (the harness with analytic fields evaluates them on the ghosts directly)
void haloExchangeInputs(svof_handle* h)
{
    if (!h->halo.active) return;
    const svof_handle::Plan& C = h->halo.cells;
    if (C.nSend) LAUNCH(h, k_gather_comps, cdiv(3LL * C.nSend, 256), 256, C.sendIdx, C.nSend, 3, h->U, C.sendBuf);
    planSendRecv(h, C, 3);
    if (C.nRecv) LAUNCH(h, k_scatter_comps, cdiv(3LL * C.nRecv, 256), 256, C.recvIdx, C.nRecv, 3, C.recvBuf, (const double*)nullptr, h->U);
    if (h->halo.facesReady) {
        const svof_handle::Plan& F = h->halo.faces;
        if (F.nSend) LAUNCH(h, k_gather_comps, cdiv(F.nSend, 256), 256, F.sendIdx, F.nSend, 1, h->phi, F.sendBuf);
        planSendRecv(h, F, 1);
        if (F.nRecv) LAUNCH(h, k_scatter_comps, cdiv(F.nRecv, 256), 256, F.recvIdx, F.nRecv, 1, F.recvBuf, F.recvSign, h->phi);
    }
}

// Build the exchange plan of a set of labelled entities (cells or faces): every rank lists the global labels of the
// entries it does not own, per owner; the owners map the labels they are asked for to their local entries.
void buildPlan(svof_handle* h, int n, const int32_t* global, const int32_t* ownerRank, const int32_t* flip, int maxComps,
               svof_handle::Plan& P, std::vector<int>* ownedOut)
{
    NcclApi& N = ncclApi();
    const int W = h->world, me = h->rank;
    std::vector<std::vector<int>> needLocal(W), needGlobal(W);
    std::vector<std::pair<int, int>> mine;   // (global, local) of the entries this rank owns
    for (int c = 0; c < n; ++c) {
        const int r = ownerRank[c];
        if (r < 0 || r >= W) throw std::invalid_argument("svof_halo_setup: owner rank out of range");
        if (r == me) { mine.emplace_back(global[c], c); if (ownedOut) ownedOut->push_back(c); }
        else { needLocal[r].push_back(c); needGlobal[r].push_back(global[c]); }
    }
    std::sort(mine.begin(), mine.end());
    P.peers.clear(); P.sendOff.assign(1, 0); P.recvOff.assign(1, 0);
    std::vector<int> sendIdx, recvIdx;
    std::vector<double> recvSign;
    if (W > 1) {
        std::vector<int> myCounts(W), all((size_t)W * W);
        for (int r = 0; r < W; ++r) myCounts[r] = (int)needLocal[r].size();
        int* dMine = dupload(h, myCounts.data(), myCounts.size());
        int* dAll = dalloc<int>(h, (size_t)W * W);
        NK(N.AllGather(dMine, dAll, (size_t)W, ncclInt32, h->nccl, h->stream));
        CK(cudaMemcpyAsync(all.data(), dAll, sizeof(int) * W * W, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        std::vector<int*> dWant(W, nullptr), dNeed(W, nullptr);
        std::vector<int> wantCount(W, 0);
        for (int r = 0; r < W; ++r) {
            wantCount[r] = (r == me) ? 0 : all[(size_t)r * W + me];
            if (wantCount[r]) dWant[r] = dalloc<int>(h, wantCount[r]);
            if (!needGlobal[r].empty()) dNeed[r] = dupload(h, needGlobal[r].data(), needGlobal[r].size());
        }
        NK(N.GroupStart());
        for (int r = 0; r < W; ++r) {
            if (r == me) continue;
            if (!needGlobal[r].empty()) NK(N.Send(dNeed[r], needGlobal[r].size(), ncclInt32, r, h->nccl, h->stream));
            if (wantCount[r]) NK(N.Recv(dWant[r], (size_t)wantCount[r], ncclInt32, r, h->nccl, h->stream));
        }
        NK(N.GroupEnd());
        CK(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < W; ++r) {
            if (r == me || (wantCount[r] == 0 && needLocal[r].empty())) continue;
            P.peers.push_back(r);
            std::vector<int> want(wantCount[r]);
            if (wantCount[r]) CK(cudaMemcpy(want.data(), dWant[r], sizeof(int) * wantCount[r], cudaMemcpyDeviceToHost));
            for (int g : want) {
                auto it = std::lower_bound(mine.begin(), mine.end(), std::make_pair(g, -1));
                if (it == mine.end() || it->first != g) throw std::invalid_argument("svof_halo_setup: a peer asks for an entry this rank does not own");
                sendIdx.push_back(it->second);
            }
            for (int c : needLocal[r]) {
                recvIdx.push_back(c);
                recvSign.push_back((flip && flip[c]) ? -1.0 : 1.0);
            }
            P.sendOff.push_back((int)sendIdx.size());
            P.recvOff.push_back((int)recvIdx.size());
        }
    }
    P.nSend = (int)sendIdx.size();
    P.nRecv = (int)recvIdx.size();
    P.sendIdx = dupload(h, sendIdx.data(), sendIdx.size());
    P.recvIdx = dupload(h, recvIdx.data(), recvIdx.size());
    P.sendBuf = dalloc<double>(h, sendIdx.size() * (size_t)maxComps);
    P.recvBuf = dalloc<double>(h, recvIdx.size() * (size_t)maxComps);
    P.recvSign = flip ? dupload(h, recvSign.data(), recvSign.size()) : nullptr;
    CK(cudaStreamSynchronize(h->stream));
}

void alphaBC(svof_handle* h)
{
    if (h->nBF > 0)
        LAUNCH(h, k_alpha_bc, cdiv(h->nBF, 256), 256, h->md, h->dPatches, h->bPatch, h->alphaBuf[h->cur], h->phi, h->alphaBBuf[h->cb]);
}

void fetchCtl(svof_handle* h);
// reconstruction::calcInterfaceNFromIsoRDF (reconstruction.C:196-405), driven from the host: the convergence test sums the
// residuals of the mixed cells in list order (avgRes, avgNormRes), so they are read back and summed here each iteration.
void rdfNormals(svof_handle* h, double* alpha)
{
    const MeshDev& d = h->md;
    cudaStream_t s = h->stream;
    if (h->capturing) throw std::runtime_error("isoRDF cannot be captured into a CUDA graph (host-side convergence test)");
    const int g128 = sparseGrid(h, 128);
    fetchCtl(h);
    const int nMixed = h->hctl->nMixed;
    h->rdfIterations = 0;
    if (nMixed == 0) return;
    if (!h->rdf) {
        h->rdf = dalloc<double>(h, h->nC);
        h->rdfB = dalloc<double>(h, std::max(h->nBF, 1));
        h->rdfBits = dalloc<unsigned int>(h, h->nWords + 1);
        h->capRdfList = h->nC;
        h->rdfList = dalloc<int>(h, h->capRdfList, false);
    }
    if (nMixed > h->capRdfMixed) {   // per-mixed-cell arrays grow with the interface (the old ones stay in the handle's pool)
        h->capRdfMixed = std::max(nMixed + nMixed / 2, 1 << 16);
        h->rdfNormal = dalloc<double>(h, 3 * (size_t)h->capRdfMixed, false);
        h->rdfRes = dalloc<double>(h, h->capRdfMixed, false);
        h->rdfAng = dalloc<double>(h, h->capRdfMixed, false);
        h->rdfSave = dalloc<double>(h, 8 * (size_t)h->capRdfMixed, false);
        h->rdfCoarse = dalloc<unsigned char>(h, h->capRdfMixed, false);
    }
    // zone = mixed cells + point neighbours; a fresh RDF per call
    if (h->nRdfPrev) LAUNCH(h, k_rdf_reset, g128, 128, d, h->rdfList, h->nRdfPrev, h->rdfBits, h->rdf, h->rdfB, 1);
    CK(cudaMemsetAsync(&h->ctl->nRdf, 0, sizeof(int), s));
    LAUNCH(h, k_rdf_mark, g128, 128, d, h->mixedCells, h->ctl, h->rdfBits, h->rdfList, h->capRdfList);
    fetchCtl(h);
    const int nRdf = std::min(h->hctl->nRdf, h->capRdfList);
    h->nRdfPrev = nRdf;
    LAUNCH(h, k_rdf_reset, g128, 128, d, h->rdfList, nRdf, h->rdfBits, h->rdf, h->rdfB, 0);
    LAUNCH(h, k_rdf_grad, g128, 128, d, h->mixedCells, h->ctl, alpha, h->alphaBBuf[h->cb], h->sp, h->rdfNormal);   // :219
    h->hRdfCoarse.assign(nMixed, 0);
    h->hRdfRes.resize(nMixed);
    h->hRdfAng.resize(nMixed);
    CK(cudaMemsetAsync(h->rdfCoarse, 0, nMixed, s));
    const double tol = h->prm.rdf_tol, relTol = h->prm.rdf_rel_tol;
    const int iterations = h->prm.rdf_iterations;
    for (int iter = 0; iter < iterations; ++iter) {
        h->rdfIterations++;
        LAUNCH(h, k_rdf_set_normals, g128, 128, h->mixedCells, h->ctl, h->rdfNormal, h->iN, h->rdfCoarse, h->cellStatus, h->iD, h->iC, h->iS,
               h->rdfSave);
        CK(cudaMemsetAsync(&h->ctl->plicNext, 0, sizeof(int), s));   // the batch counter of the persistent plane-positioning kernel
        GEO(h, plic, s, h->plicCtas, d, h->mixedCells, h->ctl, alpha, h->iN, h->sp.split, h->cellStatus, h->iD, h->iC, h->iS);
        LAUNCH(h, k_rdf_restore, g128, 128, h->mixedCells, h->ctl, h->rdfCoarse, h->cellStatus, h->iD, h->iC, h->iS, h->rdfSave);
        LAUNCH(h, k_rdf_construct, g128, 128, d, h->rdfList, h->ctl, h->iN, h->iC, h->rdf, h->rdfB);
        LAUNCH(h, k_rdf_grad, g128, 128, d, h->mixedCells, h->ctl, h->rdf, h->rdfB, h->sp, h->rdfNormal);                // :268
        LAUNCH(h, k_rdf_residual, g128, 128, d, h->mixedCells, h->ctl, h->iN, h->rdfNormal, h->rdfRes, h->rdfAng);
        CK(cudaMemcpyAsync(h->hRdfRes.data(), h->rdfRes, sizeof(double) * nMixed, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->hRdfAng.data(), h->rdfAng, sizeof(double) * nMixed, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        int resCounter = 0;
        double avgRes = 0, avgNormRes = 0;
        bool coarseChanged = false;
        for (int i = 0; i < nMixed; ++i) {
            const double normalRes = h->hRdfRes[i], avgA = h->hRdfAng[i];
            if (avgA > 0.26 && iter > 0) {   // 15 deg
                if (!h->hRdfCoarse[i]) coarseChanged = true;
                h->hRdfCoarse[i] = 1;
            } else {
                avgRes += normalRes;
                double normRes = 0;
                const double discreteError = 0.01 * (avgA * avgA);
                if (discreteError != 0) normRes = normalRes / std::max(discreteError, tol);
                else normRes = normalRes / tol;
                avgNormRes += normRes;
                resCounter++;
            }
        }
        if (resCounter == 0) {
            resCounter = 1;
            avgRes = 0;
            avgNormRes = 0;
        }
        if (((avgNormRes / resCounter < relTol || avgRes / resCounter < tol) && iter >= 1) || iter + 1 == iterations) break;
        if (coarseChanged) CK(cudaMemcpyAsync(h->rdfCoarse, h->hRdfCoarse.data(), nMixed, cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemsetAsync(&h->ctl->plicNext, 0, sizeof(int), s));   // for the plane positioning of reconstruct() that follows
}

void doReconstruct(svof_handle* h)
{
    const MeshDev& d = h->md;
    cudaStream_t s = h->stream;
    double* alpha = h->alphaBuf[h->cur];
    const int g128 = sparseGrid(h, 128);
    // A1: the front -- sparse clears of the previous step, mixed-cell list in ascending order, near sets (3 launches)
    if (h->interfaceDense) {   // mapped fields from svof_set_interface: the reference's dense zero-fill (reconstruction.C:636-641), once
        CK(cudaMemsetAsync(h->iN, 0, sizeof(double) * 3 * h->nC, s));
        CK(cudaMemsetAsync(h->iD, 0, sizeof(double) * h->nC, s));
        h->interfaceDense = false;
    }
    if (!h->bitsValid) LAUNCH(h, k_mixed_bits, cdiv(h->nC, 256), 256, alpha, h->nC, h->prm.mixed_cell_tol, h->mixedBits);
    if (h->calcBits) LAUNCH(h, k_and_bits, cdiv(h->nWords, 256), 256, h->mixedBits, h->calcBits, h->nWords);   // reconstruction.C:649-662
    LAUNCH(h, k_front_count, h->nScanBlocks, SV_SCAN_WORDS, h->mixedBits, h->nWords, h->blockSums, h->mixedCells, h->near2List, h->capNear,
           h->ctl, h->iN, h->iD, h->iC, h->iS, h->cellSlot, h->near1, h->near2);
    LAUNCH(h, k_front_scan, 1, 1024, h->blockSums, h->nScanBlocks, h->ctl, h->capMixed);
    LAUNCH(h, k_front_write, h->nScanBlocks, SV_SCAN_WORDS, d, h->mixedBits, h->nWords, h->blockSums, h->capMixed, h->mixedCells,
           h->cellStatus, h->cellSlot, h->ctl, h->near1, h->near2, h->near2List, h->capNear);
    h->epochBumps++;
    if (h->zcDenseEarly) CK(cudaEventRecord(h->evFront, s));   // zero-copy host step: list and near sets are there
    // the streaming kernel of the coming advect() only needs alpha.oldTime, phi and the near2 bitmap: with the two-stream
    // schedule it may start from here (fork 1) or once the plane-positioning kernel has been issued (fork 2)
    if (h->overlap && h->forkAt == 1) CK(cudaEventRecord(h->evNear, s));
    h->inputsAfterNear = false;
    h->freshRecon = true;
    // A2: LS normals; A3-A5: plane positions
    if (h->prm.orientation_method == SVOF_ORIENT_ISO_RDF)
        rdfNormals(h, alpha);
    else if (h->prm.orientation_method == SVOF_ORIENT_ALPHA_GRAD)
        LAUNCH(h, k_alpha_grad_normals, g128, 128, d, h->mixedCells, h->ctl, alpha, h->alphaBBuf[h->cb], h->iN, h->prm.alpha_grad_scheme == 1);
    else
        LAUNCH(h, k_ls_normals, g128, 128, d, h->mixedCells, h->ctl, alpha, h->alphaBBuf[h->cb], h->sp, h->iN);
    GEO(h, plic, s, h->plicCtas, d, h->mixedCells, h->ctl, alpha, h->iN, h->sp.split, h->cellStatus, h->iD, h->iC, h->iS);
    if (h->overlap && h->forkAt != 1) CK(cudaEventRecord(h->evNear, s));
    if (h->overlap) CK(cudaEventRecord(h->evPlic, s));
    h->bitsValid = false;  // consumed
}

void doAdvect(svof_handle* h, double dt, const double* dSp, const double* dSu)
{
    const MeshDev& d = h->md;
    const int g128 = sparseGrid(h, 128);
    double* aOld = h->alphaBuf[h->cur];
    double* aNew = h->alphaBuf[h->cur ^ 1];
    const double rDt = 1.0 / dt;
    cudaStream_t sS = h->stream, sD = h->overlap ? h->streamD : h->stream;

    // ---- streaming kernel: depends only on alpha.oldTime, phi and the near2 bitmap.  With the two-stream schedule it runs
    //      on its own low-priority stream beside the interface kernels and is joined before k_near_finalize.
    if (!h->freshRecon) LAUNCH(h, k_ctl_reset_dense, 1, 1, h->ctl);  // advect() without a new reconstruct()
    if (h->overlap) {
        if (h->zcDenseEarly) {
            // zero-copy host step: phi was pulled on this very stream and U is not read here -- only the near sets are awaited
            CK(cudaStreamWaitEvent(sD, h->evFront, 0));
        } else if (h->inputsAfterNear || dSp || dSu || !h->freshRecon) {
            CK(cudaEventRecord(h->evInputs, sS));
            CK(cudaStreamWaitEvent(sD, h->evInputs, 0));
        } else {
            CK(cudaStreamWaitEvent(sD, h->evNear, 0));
        }
    }
    if (h->prof) profBegin(h, "k_dense_update", sD);
    EventPair* ed = h->capturing ? nullptr : &beginTimedOn(h, 2, sD);
    {
        const int nTiles = cdiv(h->nC, 256);
        const int grid = (h->denseCtas > 0) ? std::min(nTiles, h->denseCtas * h->sms) : nTiles;
        if (h->useStaged)
            k_dense_update_staged<<<nTiles, 256, h->dstageSmem, sD>>>(d, h->dstage, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi,
                                                                    h->near2, h->mixedBits, dt, rDt, dSp, dSu, h->sp, h->ctl);
        else if (h->denseV4 && h->dsliced.enabled)
            k_dense_update4<<<nTiles, 256, 0, sD>>>(d, h->dsliced, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi, h->near2, h->mixedBits,
                                                    dt, rDt, dSp, dSu, h->sp, h->ctl);
        else if (h->denseV3)
            k_dense_update3<<<nTiles, 256, 0, sD>>>(d, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi, h->near2, h->mixedBits, dt, rDt,
                                                    dSp, dSu, h->sp, h->ctl);
        else if (h->dfast.enabled)
            k_dense_update2<<<grid, 256, 0, sD>>>(d, h->dfast, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi, h->near2, h->mixedBits,
                                                  dt, rDt, dSp, dSu, h->sp, h->ctl, nTiles);
        else if (h->denseCtas > 0) {   // capped grid walking the tiles: a fixed share of every SM (see k_dense_update_capped)
            const int thr = (h->denseThreads == 128) ? 128 : 256;
            const int tiles = cdiv(h->nC, thr);
            int tile0 = 0;
            if (h->overlap && h->denseSplit > 0 && h->freshRecon && !(h->inputsAfterNear || dSp || dSu)) {
                // "dense_split" percent of the tiles at full occupancy right after the near sets (beside the normals kernel,
                // until the plane-positioning kernel takes the registers), the rest capped once plane positioning is done
                tile0 = (int)((long long)tiles * std::min(h->denseSplit, 100) / 100);
                if (tile0 > 0)
                    k_dense_update_capped<<<tile0, thr, 0, sD>>>(d, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi, h->near2,
                                                                 h->mixedBits, dt, rDt, dSp, dSu, h->sp, h->ctl, 0, tile0);
                CK(cudaStreamWaitEvent(sD, h->evPlic, 0));
                h->launches++;
            }
            if (tile0 < tiles)
                k_dense_update_capped<<<std::min(tiles - tile0, h->denseCtas * h->sms), thr, 0, sD>>>(
                    d, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi, h->near2, h->mixedBits, dt, rDt, dSp, dSu, h->sp, h->ctl, tile0,
                    tiles);
        } else if (h->denseL2) {   // L2 evict-first hints on what the pass streams (see DenseMem)
#define DENSE_EF(E) k_dense_update_hint<E><<<nTiles, 256, 0, sD>>>(d, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi, h->near2, \
                                                               h->mixedBits, dt, rDt, dSp, dSu, h->sp, h->ctl)
            if (h->denseL2 == 1) DENSE_EF(1);
            else if (h->denseL2 == 2) DENSE_EF(2);
            else DENSE_EF(3);
#undef DENSE_EF
        }
        else
            k_dense_update<<<nTiles, 256, 0, sD>>>(d, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->alphaPhi, h->near2, h->mixedBits, dt, rDt,
                                                   dSp, dSu, h->sp, h->ctl);
    }
    h->launches++;
    if (ed) endTimedOn(h, *ed, sD);
    if (h->prof) profEnd(h, sD);
    if (h->overlap) CK(cudaEventRecord(h->evDense, sD));

    // ---- sparse chain
    if (!h->freshRecon) {   // advect() without a new reconstruct(): the resets k_front_scan would have done
        LAUNCH(h, k_ctl_reset_advect, 1, 1, h->ctl);
        h->epochBumps++;
    }
    // A7-A9: geometric fluxes on the downwind faces of cut cells
    if (h->un0Group)
        LAUNCH(h, k_un0_group, g128, 128, d, h->mixedCells, h->cellStatus, h->ctl, h->iN, h->iC, h->U, h->Ub, h->phi, h->Un0, h->work,
               h->capWork);
    else
        LAUNCH(h, k_un0_worklist, g128, 128, d, h->mixedCells, h->cellStatus, h->ctl, h->iN, h->iC, h->U, h->Ub, h->phi, h->Un0, h->work,
               h->capWork);
    GEO(h, faceFlux, sS, g128, d, h->work, h->ctl, h->mixedCells, h->iN, h->iD, h->Un0, h->phi, dt, h->dVfGeo);
    // A10 for the near2 cells
    LAUNCH(h, k_near_update, g128, 128, d, h->near2List, h->near1, h->ctl, aOld, aNew, h->phi, h->alphaBBuf[h->cb], h->cellSlot,
           h->cellStatus, h->dVfGeo, h->dVf, dt, rDt, dSp, dSu, h->oobList[0], h->oobState);
    // A11: conservative bounding sweeps, each proportional to the number of out-of-bounds cells
    const int gB = std::max(1, h->sms / 2);
    for (int sidx = 0; sidx < h->sp.nAlphaBounds; ++sidx) {
#define BOUND_SWEEP(MB)                                                                                                       \
    do {                                                                                                                     \
        LAUNCH(h, k_bound_deps<MB>, gB, 128, d, h->ctl, sidx, h->oobList[sidx & 1], h->oobState, aNew, aOld, h->phi, h->dVf, dSp, \
               dSu, h->bs, h->depInit, h->depLeft, h->oobIdx, (CellBound<MB>*)h->boundRecs, h->capRec, h->affList, h->boundPhiBits, h->boundPhiHost); \
        LAUNCH(h, k_bound_run<MB>, 16 * h->sms, 64, h->ctl, sidx, h->oobList[sidx & 1], h->oobState, h->bs, h->depInit, h->depLeft, h->oobIdx, \
               (const CellBound<MB>*)h->boundRecs, h->capRec, dt, rDt);                                                      \
        LAUNCH(h, k_bound_apply<MB>, gB, 128, d, h->ctl, sidx, h->affList, h->near1, aNew, h->dVf, h->bs,               \
               h->oobList[(sidx + 1) & 1], h->oobState, h->boundPhiBits);                                                                     \
    } while (0)
        if (h->maxCF <= 8 && h->boundLanes) {   // eight faces in eight lanes of the chain walker's warp
            LAUNCH(h, k_bound_deps<8>, gB, 128, d, h->ctl, sidx, h->oobList[sidx & 1], h->oobState, aNew, aOld, h->phi, h->dVf, dSp, dSu, h->bs,
                   h->depInit, h->depLeft, h->oobIdx, (CellBound<8>*)h->boundRecs, h->capRec, h->affList, h->boundPhiBits, h->boundPhiHost);
            LAUNCH(h, k_bound_run8, 16 * h->sms, 64, h->ctl, sidx, h->oobList[sidx & 1], h->oobState, h->bs, h->depInit, h->depLeft, h->oobIdx,
                   (const CellBound<8>*)h->boundRecs, h->capRec, dt, rDt);
            LAUNCH(h, k_bound_apply<8>, gB, 128, d, h->ctl, sidx, h->affList, h->near1, aNew, h->dVf, h->bs, h->oobList[(sidx + 1) & 1],
                   h->oobState, h->boundPhiBits);
        } else if (h->maxCF <= 8) BOUND_SWEEP(8);
        else if (h->maxCF <= 16) BOUND_SWEEP(16);
        else BOUND_SWEEP(64);
    }
    // join: the finalize kernel ORs into the bitmap words the streaming kernel wrote
    if (h->overlap) CK(cudaStreamWaitEvent(sS, h->evDense, 0));
    // A12: snap/clip + alphaPhi for near2; boundary values of the new field
    LAUNCH(h, k_near_finalize, g128, 128, d, h->near2List, h->ctl, aNew, h->dVf, h->alphaPhi, h->mixedBits, dt, h->sp, h->oobState);
    h->cur ^= 1;
    h->cb ^= 1;  // keep the patch values alpha.oldTime() was advected with (for dVf materialisation)
    h->bitsValid = true;
    if (!(h->capturing && h->halo.active)) {  // decomposed runs: the ghost refresh (NCCL) stays outside the captured graph
        haloExchange(h);
        alphaBC(h);
    }
    h->advected = true;
    h->freshRecon = false;
    h->lastDt = dt;
    ++h->advectCount;
    if (h->epochBumps >= (1 << 25)) {  // tags would wrap: clear the tag arrays
        CK(cudaMemsetAsync(h->bs.tagV, 0, sizeof(int) * h->nF, sS));
        CK(cudaMemsetAsync(h->bs.tagR, 0, sizeof(int) * h->nF, sS));
        CK(cudaMemsetAsync(h->bs.affStamp, 0, sizeof(int) * h->nC, sS));
        CK(cudaMemsetAsync(&h->ctl->epoch, 0, sizeof(int), sS));
        h->epochBumps = 0;
    }
}

// every run-time option that changes which kernels / grids a captured step contains
int schedKey(const svof_handle* h)
{
    return (((h->overlap * 10 + h->forkAt) * 100 + h->plicCtas) * 100 + h->denseCtas) * 8 + (h->denseThreads == 128 ? 1 : 0) + 2 * h->denseL2 + 1000000 * h->denseSplit;
}

void fetchCtl(svof_handle* h)
{
    CK(cudaMemcpyAsync(h->hctl, h->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
}

int fail(svof_handle* h, int code, const char* what)
{
    if (h) h->err = what;
    return code;
}

#define API_BEGIN try {
#define API_END(h)                                           \
    }                                                        \
    catch (const std::invalid_argument& e) { return fail(h, SVOF_ERR_INVALID_ARG, e.what()); } \
    catch (const std::length_error& e) { return fail(h, SVOF_ERR_CAPACITY, e.what()); }        \
    catch (const CommError& e) { return fail(h, SVOF_ERR_COMM, e.what()); }                    \
    catch (const std::exception& e) { return fail(h, SVOF_ERR_CUDA, e.what()); }

bool parseBool(const char* v, int32_t* out)
{
    static const char* T[] = {"true", "on", "yes", "y", "t", "1"};
    static const char* F[] = {"false", "off", "no", "n", "f", "0", "none"};
    for (const char* s : T) if (!strcmp(v, s)) { *out = 1; return true; }
    for (const char* s : F) if (!strcmp(v, s)) { *out = 0; return true; }
    return false;
}

int deviceErr(svof_handle* h);
int checkDeviceErr(svof_handle* h)
{
    fetchCtl(h);
    return h->hctl->err ? deviceErr(h) : (int)SVOF_OK;
}
// h->hctl holds a fresh copy of the control block with err != 0
int deviceErr(svof_handle* h)
{
    {
        char b[256];
        snprintf(b, sizeof(b), "device capacity flag 0x%x (1 face verts, 2 cell faces, 4 cell points, 8 interface points, "
                 "16 LS stencil, 32 work list)", h->hctl->err);
        h->err = b;
        return SVOF_ERR_CAPACITY;
    }
    return SVOF_OK;
}

}  // namespace

extern "C" {

int svof_params_default(svof_params* p)
{
    if (!p) return SVOF_ERR_INVALID_ARG;
    memset(p, 0, sizeof(*p));
    p->mixed_cell_tol = 1e-8;
    p->snap_tol = 0.0;
    p->rdf_tol = 1e-6;
    p->rdf_rel_tol = 0.1;
    p->n_alpha_bounds = 10;
    p->clip = 1;
    p->alpha_grad_scheme = 0;
    p->orientation_method = SVOF_ORIENT_ISO_ALPHA_GRAD;
    p->rdf_iterations = 5;
    return SVOF_OK;
}

int svof_params_set(svof_params* p, const char* key, const char* value)
{
    if (!p || !key || !value) return SVOF_ERR_INVALID_ARG;
    std::string k(key), v(value);
    while (!v.empty() && (v.back() == ';' || v.back() == ' ')) v.pop_back();
    char* end = nullptr;
    auto num = [&](double* out) {
        *out = strtod(v.c_str(), &end);
        return end != v.c_str() && *end == '\0';
    };
    double dv;
    int32_t b;
    if (k == "mixedCellTol") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; p->mixed_cell_tol = dv; p->mixed_cell_tol_set = 1; return SVOF_OK; }
    if (k == "surfCellTol") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; if (!p->mixed_cell_tol_set) p->mixed_cell_tol = dv; return SVOF_OK; }
    if (k == "isoFaceTol") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; p->iso_face_tol = dv; return SVOF_OK; }
    if (k == "snapTol") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; p->snap_tol = dv; return SVOF_OK; }
    if (k == "tol") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; p->rdf_tol = dv; return SVOF_OK; }
    if (k == "relTol") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; p->rdf_rel_tol = dv; return SVOF_OK; }
    if (k == "nAlphaBounds") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; p->n_alpha_bounds = (int32_t)dv; return SVOF_OK; }
    if (k == "iterations") { if (!num(&dv)) return SVOF_ERR_INVALID_ARG; p->rdf_iterations = (int32_t)dv; return SVOF_OK; }
    if (k == "clip") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->clip = b; return SVOF_OK; }
    if (k == "splitWarpedFace") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->split_warped_face = b; return SVOF_OK; }
    if (k == "mapAlphaField") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->map_alpha_field = b; return SVOF_OK; }
    if (k == "writePlicFields") { if (!parseBool(v.c_str(), &b)) return SVOF_ERR_INVALID_ARG; p->write_plic_fields = b; return SVOF_OK; }
    if (k == "orientationMethod") {
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
    if (k == "nAlphaSubCycles" || k == "cAlpha" || k == "period" || k == "reverseTime" || k == "nAlphaCorr") return SVOF_OK;
    return SVOF_ERR_INVALID_ARG;
}

int svof_create(const svof_mesh* mesh, const svof_params* params, const svof_comm* comm, svof_handle** out)
{
    if (!mesh || !params || !out) { g_createError = "svof_create: null argument"; return SVOF_ERR_INVALID_ARG; }
    if (comm && comm->world_size > 1) {
        if (comm->rank < 0 || comm->rank >= comm->world_size) { g_createError = "svof_comm: rank out of range"; return SVOF_ERR_INVALID_ARG; }
        if (!comm->nccl_unique_id) { g_createError = "svof_comm: world_size > 1 needs nccl_unique_id (svof_comm_unique_id on one rank, broadcast)"; return SVOF_ERR_INVALID_ARG; }
        if (!ncclApi().ok) { g_createError = "NCCL unavailable: " + ncclApi().why; return SVOF_ERR_COMM; }
    }
    if (params->orientation_method == SVOF_ORIENT_ISO_RDF && comm && comm->world_size > 1) {
        g_createError = "orientationMethod isoRDF in a decomposed run: the convergence sums (reconstruction.C:369-371) are not reduced across ranks yet";
        return SVOF_ERR_UNSUPPORTED;
    }
    if (params->n_alpha_bounds > SV_MAX_SWEEPS) { g_createError = "nAlphaBounds exceeds 32"; return SVOF_ERR_INVALID_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_createError = "no CUDA device: this library has no CPU path";
        (void)cudaGetLastError();
        return SVOF_ERR_CUDA;
    }
    svof_handle* h = new svof_handle;
    int rc = SVOF_OK;
    try {
        h->device = (comm && comm->device >= 0) ? comm->device : ((comm ? comm->rank : 0) % ndev);
        CK(cudaSetDevice(h->device));
        // the sparse chain is the latency-critical one: give its CTAs priority over the streaming kernel's
        int prLo = 0, prHi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
        CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prHi));
        CK(cudaStreamCreateWithPriority(&h->streamD, cudaStreamNonBlocking, prLo));
        CK(cudaStreamCreateWithPriority(&h->streamU, cudaStreamNonBlocking, prHi));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, h->device));
        h->sms = prop.multiProcessorCount;
        h->overlap = getenv("SVOF_OVERLAP") ? atoi(getenv("SVOF_OVERLAP")) : 1;  // default on (round 2: 1.173 vs 1.206 ms/step at 256^3)
        if (getenv("SVOF_FORK")) { h->forkAt = atoi(getenv("SVOF_FORK")); h->schedUser = true; }
        if (getenv("SVOF_OVERLAP") || getenv("SVOF_DENSE_CTAS")) h->schedUser = true;
        if (getenv("SVOF_SCHED_AUTO")) h->tuneMode = atoi(getenv("SVOF_SCHED_AUTO")) != 0;
        if (getenv("SVOF_BOUND_LANES")) h->boundLanes = atoi(getenv("SVOF_BOUND_LANES")) != 0;
        if (getenv("SVOF_UN0") && !strcmp(getenv("SVOF_UN0"), "group")) h->un0Group = true;   // measured 84 us against 77 us: opt-in
        if (getenv("SVOF_DENSE_CTAS")) h->denseCtas = atoi(getenv("SVOF_DENSE_CTAS"));
        h->prof = getenv("SVOF_PROFILE") && atoi(getenv("SVOF_PROFILE")) > 0;
        h->prm = *params;
        h->sp.mixedTol = params->mixed_cell_tol;
        h->sp.snapTol = params->snap_tol;
        h->sp.clip = params->clip;
        h->sp.nAlphaBounds = params->n_alpha_bounds;
        h->sp.split = params->split_warped_face;
        buildMesh(h, *mesh);
        allocFields(h);
        CK(cudaStreamSynchronize(h->stream));
        if (comm && comm->world_size > 1) {
            h->rank = comm->rank;
            h->world = comm->world_size;
            ncclUniqueId id;
            memcpy(&id, comm->nccl_unique_id, sizeof(id));
            NK(ncclApi().CommInitRank(&h->nccl, h->world, id, h->rank));
        }
    } catch (const Unsupported& e) { g_createError = e.what(); rc = SVOF_ERR_UNSUPPORTED; }
    catch (const std::invalid_argument& e) { g_createError = e.what(); rc = SVOF_ERR_BAD_MESH; }
    catch (const std::length_error& e) { g_createError = e.what(); rc = SVOF_ERR_CAPACITY; }
    catch (const CommError& e) { g_createError = e.what(); rc = SVOF_ERR_COMM; }
    catch (const std::exception& e) { g_createError = e.what(); rc = SVOF_ERR_CUDA; }
    if (rc) {
        svof_destroy(h);
        return rc;
    }
    *out = h;
    return SVOF_OK;
}

int svof_destroy(svof_handle* h)
{
    if (!h) return SVOF_OK;
    cudaSetDevice(h->device);
    if (h->prof) profPrint(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->streamD) cudaStreamSynchronize(h->streamD);
    if (h->nccl && ncclApi().ok) ncclApi().CommDestroy(h->nccl);
    for (void* p : h->allocs) cudaFree(p);
    if (h->hctl) cudaFreeHost(h->hctl);
    if (h->hpartial) cudaFreeHost(h->hpartial);
    if (h->hUList) cudaFreeHost(h->hUList);
    if (h->hUPacked) cudaFreeHost(h->hUPacked);
    if (h->hIdx) cudaFreeHost(h->hIdx);
    if (h->hVal) cudaFreeHost(h->hVal);
    for (int i = 0; i < 2; ++i) {
        if (h->hPhiBits2[i]) cudaFreeHost(h->hPhiBits2[i]);
        if (h->hPhiBlockOff2[i]) cudaFreeHost(h->hPhiBlockOff2[i]);
    }
    if (h->hPhiPacked) cudaFreeHost(h->hPhiPacked);
    if (h->evBits) cudaEventDestroy(h->evBits);
    delete h->pool;
    for (EventPair& e : h->events) {
        if (e.a) cudaEventDestroy(e.a);
        if (e.b) cudaEventDestroy(e.b);
    }
    for (int i = 0; i < 8; ++i) if (h->marks[i]) cudaEventDestroy(h->marks[i]);
    for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->evNear) cudaEventDestroy(h->evNear);
    if (h->evPlic) cudaEventDestroy(h->evPlic);
    if (h->evDense) cudaEventDestroy(h->evDense);
    if (h->evInputs) cudaEventDestroy(h->evInputs);
    if (h->evCopy) cudaEventDestroy(h->evCopy);
    if (h->evFront) cudaEventDestroy(h->evFront);
    if (h->evU) cudaEventDestroy(h->evU);
    if (h->evPush) cudaEventDestroy(h->evPush);
    if (h->evPhiNear) cudaEventDestroy(h->evPhiNear);
    for (int i = 0; i < 4; ++i) if (h->evT[i]) cudaEventDestroy(h->evT[i]);
    if (h->streamU) { cudaStreamSynchronize(h->streamU); cudaStreamDestroy(h->streamU); }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->streamD) cudaStreamDestroy(h->streamD);
    delete h;
    return SVOF_OK;
}

const char* svof_last_error(const svof_handle* h) { return h ? h->err.c_str() : g_createError.c_str(); }

int svof_set_alpha(svof_handle* h, const double* alpha)
{
    if (!h || !alpha) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->alphaBuf[h->cur], alpha, sizeof(double) * h->nC, cudaMemcpyHostToDevice, h->stream));
    alphaBC(h);
    CK(cudaStreamSynchronize(h->stream));
    h->haveAlpha = true;
    h->bitsValid = false;
    h->advected = false;
    h->hostAlphaSynced = nullptr; h->phiBitsReady = false;
    return SVOF_OK;
    API_END(h)
}

int svof_set_phi(svof_handle* h, const double* phi)
{
    if (!h || !phi) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->phi, phi, sizeof(double) * h->nF, cudaMemcpyHostToDevice, h->stream));
    h->havePhi = true;
    h->phiPartial = false;
    if (h->anyInletOutlet && h->haveAlpha) alphaBC(h);   // inletOutlet patch values depend on the sign of phi
    CK(cudaStreamSynchronize(h->stream));
    return SVOF_OK;
    API_END(h)
}

int svof_set_U(svof_handle* h, const double* U, const double* Ub)
{
    if (!h || !U) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->U, U, sizeof(double) * 3 * h->nC, cudaMemcpyHostToDevice, h->stream));
    if (Ub && h->nBF) CK(cudaMemcpyAsync(h->Ub, Ub, sizeof(double) * 3 * h->nBF, cudaMemcpyHostToDevice, h->stream));
    else if (h->nBF) CK(cudaMemsetAsync(h->Ub, 0, sizeof(double) * 3 * h->nBF, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->haveU = true;
    h->uPartial = false;
    return SVOF_OK;
    API_END(h)
}

int svof_scatter_alpha_device(svof_handle* h, const int32_t* d_idx, const double* d_vals, int64_t n)
{
    if (!h || n < 0 || (n > 0 && (!d_idx || !d_vals))) return SVOF_ERR_INVALID_ARG;
    if (!h->haveAlpha) return fail(h, SVOF_ERR_STATE, "svof_scatter_alpha_device: alpha not set");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    if (n) LAUNCH(h, k_scatter_alpha, (int)cdiv(n, 256), 256, d_idx, d_vals, (long long)n, h->prm.mixed_cell_tol, h->alphaBuf[h->cur],
                  h->mixedBits, h->bitsValid ? 1 : 0);
    alphaBC(h);
    h->advected = false;
    h->hostAlphaSynced = nullptr; h->phiBitsReady = false;
    return SVOF_OK;
    API_END(h)
}


// ---- changing meshes ------------------------------------------------------------------------------------
int svof_update_points(svof_handle* h, const double* points, const double* Cf, const double* Sf, const double* C, const double* V)
{
    if (!h || !points) return SVOF_ERR_INVALID_ARG;
    if ((Cf || Sf || C || V) && !(Cf && Sf && C && V)) return fail(h, SVOF_ERR_INVALID_ARG, "svof_update_points: give all of Cf, Sf, C, V or none");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->streamD));
    CK(cudaMemcpyAsync(const_cast<double*>(h->md.points), points, sizeof(double) * 3 * h->nP, cudaMemcpyHostToDevice, h->stream));
    computeGeometry(h, Cf, Sf, C, V);   // reconstruction.C:643-647: flatness follows the mesh
    h->advected = false;
    return SVOF_OK;
    API_END(h)
}

namespace {
void releaseMeshState(svof_handle* h)
{
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->streamD) cudaStreamSynchronize(h->streamD);
    for (auto& g : h->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
    h->bytes = 0;
    auto freeHost = [](auto*& p) { if (p) cudaFreeHost(p); p = nullptr; };
    freeHost(h->hctl); freeHost(h->hpartial); freeHost(h->hUList); freeHost(h->hUPacked); freeHost(h->hIdx); freeHost(h->hVal);
    for (int i = 0; i < 2; ++i) { freeHost(h->hPhiBits2[i]); freeHost(h->hPhiBlockOff2[i]); }
    freeHost(h->hPhiPacked);
    h->phiPrevBitsValid = h->alphaPhiPrevValid = false;
    h->phiBits = nullptr; h->phiBlockOff = nullptr; h->phiPacked = nullptr; h->phiPartial = false;
    h->rdf = nullptr; h->capRdfMixed = 0; h->nRdfPrev = 0;   // isoRDF buffers are re-created for the new mesh
    h->calcBits = h->calcBitsStore = nullptr;                // cell types belong to the old mesh
    h->zcBits[0] = h->zcBits[1] = nullptr; h->zcPrevValid = false;
    harvestEvents(h, true);
    h->halo = svof_handle::Halo();
    h->dfast = DenseFast();
    h->haveAlpha = h->havePhi = h->haveU = h->bitsValid = h->advected = h->uPartial = h->phiPartial = false;
    h->freshRecon = h->inputsAfterNear = false;
    h->hostAlphaSynced = h->hostAlphaPhiSynced = nullptr; h->phiBitsReady = false;
    h->cur = h->cb = 0;
    h->epochBumps = 0;
}
}  // namespace

int svof_update_mesh(svof_handle* h, const svof_mesh* mesh)
{
    if (!h || !mesh) return SVOF_ERR_INVALID_ARG;
    int rc = SVOF_OK;
    try {
        CK(cudaSetDevice(h->device));
        releaseMeshState(h);
        buildMesh(h, *mesh);
        allocFields(h);
        CK(cudaStreamSynchronize(h->stream));
    } catch (const Unsupported& e) { h->err = e.what(); rc = SVOF_ERR_UNSUPPORTED; }
    catch (const std::invalid_argument& e) { h->err = e.what(); rc = SVOF_ERR_BAD_MESH; }
    catch (const std::length_error& e) { h->err = e.what(); rc = SVOF_ERR_CAPACITY; }
    catch (const std::exception& e) { h->err = e.what(); rc = SVOF_ERR_CUDA; }
    return rc;
}

int svof_set_interface(svof_handle* h, const double* interfaceN, const double* interfaceD)
{
    if (!h || !interfaceN || !interfaceD) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->iN, interfaceN, sizeof(double) * 3 * h->nC, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->iD, interfaceD, sizeof(double) * h->nC, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->interfaceDense = true;   // the next reconstruct() must clear the whole fields, not just the previous mixed cells
    return SVOF_OK;
    API_END(h)
}

int svof_set_cell_types(svof_handle* h, const int32_t* cell_types)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    const bool had = h->calcBits != nullptr;
    if (!cell_types) {
        h->calcBits = nullptr;   // the buffer stays with the handle's allocations; the filter is off
    } else {
        std::vector<unsigned int> bits((size_t)h->nWords, 0u);
        for (int c = 0; c < h->nC; ++c)
            if (cell_types[c] == 0) bits[(size_t)c >> 5] |= 1u << (c & 31);   // cellCellStencil::CALCULATED
        if (!h->calcBitsStore) h->calcBitsStore = dalloc<unsigned int>(h, (size_t)h->nWords);
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(h->calcBitsStore, bits.data(), sizeof(unsigned int) * (size_t)h->nWords, cudaMemcpyHostToDevice));
        h->calcBits = h->calcBitsStore;
    }
    // the launch list of reconstruct() changes with the filter: captured steps are re-captured
    if (had != (h->calcBits != nullptr))
        for (auto& g : h->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    return SVOF_OK;
    API_END(h)
}

int svof_map_alpha_field(svof_handle* h, double lower, double upper)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (!h->haveAlpha) return fail(h, SVOF_ERR_STATE, "svof_map_alpha_field: alpha not set");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    const auto t0 = std::chrono::steady_clock::now();
    double* alpha = h->alphaBuf[h->cur];
    const int v = (h->variant >= 3) ? 2 : h->variant;   // the non-split capacity variant of this mesh (reconstruction.C:768)
    if (h->prof) profBegin(h, "mapAlpha", h->stream);
    switch (v) {
        case 0: GeoLaunch<CapsHex>::mapAlpha(h->stream, h->md, h->iN, h->iD, lower, upper, alpha, h->ctl); break;
        case 1: GeoLaunch<CapsSmall>::mapAlpha(h->stream, h->md, h->iN, h->iD, lower, upper, alpha, h->ctl); break;
        default: GeoLaunch<CapsPoly>::mapAlpha(h->stream, h->md, h->iN, h->iD, lower, upper, alpha, h->ctl); break;
    }
    if (h->prof) profEnd(h, h->stream);
    h->launches++;
    alphaBC(h);                 // alpha1_.correctBoundaryConditions(); alpha1_.oldTime() = alpha1_ (the old-time buffer is the current one)
    h->bitsValid = false;
    h->advected = false;
    h->hostAlphaSynced = nullptr; h->phiBitsReady = false;
    CK(cudaStreamSynchronize(h->stream));
    h->mapTime += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return checkDeviceErr(h);
    API_END(h)
}

// ---- decomposed runs: NCCL bootstrap and the ghost-refresh plan --------------------------------------
int svof_comm_unique_id(void* id128)
{
    if (!id128) return SVOF_ERR_INVALID_ARG;
    NcclApi& N = ncclApi();
    if (!N.ok) { g_createError = "NCCL unavailable: " + N.why; return SVOF_ERR_COMM; }
    ncclUniqueId id;
    if (N.GetUniqueId(&id) != ncclSuccess) { g_createError = "ncclGetUniqueId failed"; return SVOF_ERR_COMM; }
    memcpy(id128, &id, sizeof(id));
    return SVOF_OK;
}

int svof_halo_setup(svof_handle* h, const int32_t* cell_global, const int32_t* cell_owner_rank)
{
    if (!h || !cell_global || !cell_owner_rank) return SVOF_ERR_INVALID_ARG;
    if (h->world > 1 && !h->nccl) return fail(h, SVOF_ERR_STATE, "svof_halo_setup: the handle was created without an NCCL id");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    for (int c = 1; c < h->nC; ++c)
        if (cell_global[c - 1] >= cell_global[c]) throw std::invalid_argument("svof_halo_setup: cell_global must ascend");
    std::vector<int> owned;
    buildPlan(h, h->nC, cell_global, cell_owner_rank, nullptr, 3, h->halo.cells, &owned);
    h->halo.nOwned = (int)owned.size();
    h->halo.ownedIdx = dupload(h, owned.data(), owned.size());
    CK(cudaStreamSynchronize(h->stream));
    h->halo.active = (h->world > 1);
    for (auto& g : h->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }   // the captured tail changes
    return SVOF_OK;
    API_END(h)
}

int svof_halo_setup_faces(svof_handle* h, const int32_t* face_global, const int32_t* face_owner_rank, const int32_t* face_flip)
{
    if (!h || !face_global || !face_owner_rank) return SVOF_ERR_INVALID_ARG;
    if (h->world > 1 && !h->nccl) return fail(h, SVOF_ERR_STATE, "svof_halo_setup_faces: the handle was created without an NCCL id");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    buildPlan(h, h->nF, face_global, face_owner_rank, face_flip, 1, h->halo.faces, nullptr);
    h->halo.facesReady = true;
    return SVOF_OK;
    API_END(h)
}

int svof_halo_exchange_inputs(svof_handle* h)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (!h->havePhi || !h->haveU) return fail(h, SVOF_ERR_STATE, "svof_halo_exchange_inputs: phi/U not set");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    haloExchangeInputs(h);
    h->inputsAfterNear = true;
    return SVOF_OK;
    API_END(h)
}

int svof_halo_exchange(svof_handle* h)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (!h->haveAlpha) return fail(h, SVOF_ERR_STATE, "svof_halo_exchange: alpha not set");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    haloExchange(h);
    alphaBC(h);
    h->hostAlphaSynced = nullptr; h->phiBitsReady = false;
    return SVOF_OK;
    API_END(h)
}

int svof_set_phi_device(svof_handle* h, const void* dphi)
{
    if (!h || !dphi) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->phi, dphi, sizeof(double) * h->nF, cudaMemcpyDeviceToDevice, h->stream));
    h->havePhi = true;
    h->phiPartial = false;
    h->inputsAfterNear = true;
    if (h->anyInletOutlet && h->haveAlpha) alphaBC(h);
    return SVOF_OK;
    API_END(h)
}

int svof_set_U_device(svof_handle* h, const void* dU, const void* dUb)
{
    if (!h || !dU) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->U, dU, sizeof(double) * 3 * h->nC, cudaMemcpyDeviceToDevice, h->stream));
    if (dUb && h->nBF) CK(cudaMemcpyAsync(h->Ub, dUb, sizeof(double) * 3 * h->nBF, cudaMemcpyDeviceToDevice, h->stream));
    h->haveU = true;
    h->uPartial = false;
    return SVOF_OK;
    API_END(h)
}

int svof_reconstruct(svof_handle* h)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (!h->haveAlpha) return fail(h, SVOF_ERR_STATE, "svof_reconstruct: alpha not set");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    EventPair& e = beginTimed(h, 0);
    doReconstruct(h);
    endTimed(h, e);
    CK(cudaGetLastError());
    return SVOF_OK;
    API_END(h)
}

int svof_advect(svof_handle* h, double dt, const double* Sp, const double* Su)
{
    if (!h || !(dt > 0)) return SVOF_ERR_INVALID_ARG;
    if (!h->haveAlpha || !h->havePhi || !h->haveU) return fail(h, SVOF_ERR_STATE, "svof_advect: alpha/phi/U not set");
    if (h->uPartial) return fail(h, SVOF_ERR_STATE, "svof_advect: U on the device is the sparse upload of svof_step_host; call svof_set_U first");
    if (h->phiPartial) return fail(h, SVOF_ERR_STATE, "svof_advect: phi on the device is the sparse upload of svof_step_host; call svof_set_phi first");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    const double *dSp = nullptr, *dSu = nullptr;
    if (Sp) { CK(cudaMemcpyAsync(h->Sp, Sp, sizeof(double) * h->nC, cudaMemcpyHostToDevice, h->stream)); dSp = h->Sp; }
    if (Su) { CK(cudaMemcpyAsync(h->Su, Su, sizeof(double) * h->nC, cudaMemcpyHostToDevice, h->stream)); dSu = h->Su; }
    EventPair& e = beginTimed(h, 1);
    doAdvect(h, dt, dSp, dSu);
    endTimed(h, e);
    h->hostAlphaSynced = h->hostAlphaPhiSynced = nullptr; h->phiBitsReady = false;  // the caller's buffers no longer mirror the device
    CK(cudaGetLastError());
    if (Sp || Su) CK(cudaStreamSynchronize(h->stream));  // caller's buffers may be pageable
    return SVOF_OK;
    API_END(h)
}

namespace {
constexpr int kTuneTimed = 12;   // timed steps per schedule (4 were too few next to a clock sampler: profiles/r5k)
void tuneUse(svof_handle* h, int slot)
{
    h->tuneSlot = slot;
    h->forkAt = slot ? 2 : h->fork0;
    h->denseCtas = slot ? 4 : h->dense0;
}
// an explicit schedule option: the run-time selection steps aside and hands back the schedule as configured
void schedByUser(svof_handle* h)
{
    if (!h->schedUser && h->tuneSlot != 0) tuneUse(h, 0);
    h->schedUser = true;
}
// called once per svof_step_device before the step is enqueued
void tuneAdvance(svof_handle* h)
{
    if (!h->tuneMode || h->schedUser || h->halo.active || !h->overlap || h->useStaged || h->denseV3 || h->denseV4 || h->dfast.enabled) return;
    cudaStream_t st = h->stream;
    switch (h->tunePhase) {
        case 0:   // slot 0 warming up (its graphs are captured during these steps)
            if (h->tuneSteps == 0) { h->fork0 = h->forkAt; h->dense0 = h->denseCtas; tuneUse(h, 0); }
            if (h->tuneSteps >= 6) { CK(cudaEventRecord(h->evT[0], st)); h->tunePhase = 1; h->tuneSteps = 0; }
            break;
        case 1:   // slot 0 timed
            if (h->tuneSteps >= kTuneTimed) { CK(cudaEventRecord(h->evT[1], st)); tuneUse(h, 1); h->tunePhase = 2; h->tuneSteps = 0; }
            break;
        case 2:   // slot 1 warming up
            if (h->tuneSteps >= 6) { CK(cudaEventRecord(h->evT[2], st)); h->tunePhase = 3; h->tuneSteps = 0; }
            break;
        case 3:   // slot 1 timed, then the decision (the one host wait of the selection)
            if (h->tuneSteps >= kTuneTimed) {
                CK(cudaEventRecord(h->evT[3], st));
                CK(cudaEventSynchronize(h->evT[3]));
                float a = 0, b = 0;
                CK(cudaEventElapsedTime(&a, h->evT[0], h->evT[1]));
                CK(cudaEventElapsedTime(&b, h->evT[2], h->evT[3]));
                h->tuneMs[0] = a / kTuneTimed; h->tuneMs[1] = b / kTuneTimed;
                tuneUse(h, (b < 0.98 * a) ? 1 : 0);   // the alternative has to win by 2 %
                if (getenv("SVOF_SCHED_DEBUG"))
                    fprintf(stderr, "[svof sched] default %.4f ms/step, alternative %.4f ms/step -> %s\n", h->tuneMs[0], h->tuneMs[1],
                            h->tuneSlot ? "alternative" : "default");
                h->tunePhase = 4; h->tuneSteps = 0; h->retune = false;
            }
            break;
        default:  // settled: measure again now and then (the interface grows and shrinks during a run), or on request
            if (h->retune || h->tuneSteps >= 4096) { tuneUse(h, 0); h->tunePhase = 0; h->tuneSteps = 1; h->retune = false; }
            break;
    }
}
}  // namespace

// reconstruct() + advect(dt) with device-resident inputs as ONE CUDA-graph launch.  The step is ~27 dependent launches,
// most of them a few microseconds long; replaying a captured graph removes the per-launch gaps.  Every launch size is
// host-known and every count lives on the device, so the captured graph is valid for any state of the fields; it is
// keyed by the two buffer parities and re-captured when dt changes.
int svof_step_device(svof_handle* h, double dt)
{
    if (!h || !(dt > 0)) return SVOF_ERR_INVALID_ARG;
    if (!h->haveAlpha || !h->havePhi || !h->haveU) return fail(h, SVOF_ERR_STATE, "svof_step_device: alpha/phi/U not set");
    if (h->uPartial) return fail(h, SVOF_ERR_STATE, "svof_step_device: U on the device is the sparse upload of svof_step_host; call svof_set_U first");
    if (h->phiPartial) return fail(h, SVOF_ERR_STATE, "svof_step_device: phi on the device is the sparse upload of svof_step_host; call svof_set_phi first");
    if (h->prof || h->epochBumps >= (1 << 25) - 4 || h->prm.orientation_method == SVOF_ORIENT_ISO_RDF) {
        // instrumented runs / tag wrap imminent / isoRDF (host-side convergence test): plain launches
        const int rc = svof_reconstruct(h);
        return rc ? rc : svof_advect(h, dt, nullptr, nullptr);
    }
    API_BEGIN
    CK(cudaSetDevice(h->device));
    tuneAdvance(h);
    h->tuneSteps++;
    svof_handle::StepGraph& g = h->graphs[h->tuneSlot * 8 + h->cur * 4 + h->cb * 2 + (h->bitsValid ? 1 : 0)];
    if (!g.exec || g.dt != dt || g.sched != schedKey(h)) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        // run the first step of this kind with plain launches (warms every lazily initialised launcher), then capture
        const int cur0 = h->cur, cb0 = h->cb;
        const bool bits0 = h->bitsValid;
        const int rc = svof_reconstruct(h);
        if (rc) return rc;
        const int rc2 = svof_advect(h, dt, nullptr, nullptr);
        if (rc2) return rc2;
        // capture the step that starts from (cur0, cb0) without executing it: restore the host-side state it starts from
        const int curN = h->cur, cbN = h->cb;
        const long long l0 = h->launches;
        const int ac0 = h->advectCount;
        h->cur = cur0; h->cb = cb0; h->bitsValid = bits0;
        h->capturing = true;
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
        if (ce == cudaSuccess) {
            try {
                doReconstruct(h);
                doAdvect(h, dt, nullptr, nullptr);
            } catch (...) {
                cudaStreamEndCapture(h->stream, &graph);
                if (graph) cudaGraphDestroy(graph);
                h->capturing = false;
                h->cur = curN; h->cb = cbN; h->advectCount = ac0; h->launches = l0;
                throw;
            }
            ce = cudaStreamEndCapture(h->stream, &graph);
        }
        h->capturing = false;
        g.nLaunches = h->launches - l0;
        h->cur = curN; h->cb = cbN; h->advectCount = ac0; h->launches = l0;   // the capture executed nothing
        if (ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { g.exec = nullptr; (void)cudaGetLastError(); }   // graphs unavailable: keep using plain launches
        g.dt = dt;
        g.sched = schedKey(h);
        return SVOF_OK;
    }
    CK(cudaGraphLaunch(g.exec, h->stream));
    // the host-side state transitions of doReconstruct() + doAdvect()
    h->launches += g.nLaunches;
    h->cur ^= 1;
    h->cb ^= 1;
    h->bitsValid = true;
    if (h->halo.active) {
        haloExchange(h);
        alphaBC(h);
    }
    h->advected = true;
    h->freshRecon = false;
    h->inputsAfterNear = false;
    h->lastDt = dt;
    h->advectCount++;
    h->hostAlphaSynced = h->hostAlphaPhiSynced = nullptr; h->phiBitsReady = false;
    return SVOF_OK;
    API_END(h)
}

// profile only: host wall time since the previous tick goes to bucket `name` (ticks sit after existing syncs)
inline void hostTick(svof_handle* h, const char* name)
{
    if (!h->prof) return;
    const auto now = std::chrono::steady_clock::now();
    if (name) h->hostAcc[name] += std::chrono::duration<double, std::milli>(now - h->hostT).count();
    h->hostT = now;
}


namespace {
HostPool& hostPool(svof_handle* h)
{
    if (!h->pool) h->pool = new HostPool((int)std::max(1u, std::min(std::thread::hardware_concurrency(), 8u)) - 1);
    return *h->pool;
}

// face bitmap of the current alpha + marked faces per 1024-face block, and their read-back, on stream s
void enqueuePhiBits(svof_handle* h, cudaStream_t s, int buf)
{
    CK(cudaMemsetAsync(h->phiBlockOff, 0, sizeof(int) * ((size_t)h->nPhiBlocks + 1), s));
    k_phi_need_bits<<<cdiv((long long)h->nWordsF * 32, 256), 256, 0, s>>>(h->md, h->alphaBuf[h->cur], h->phiBits, h->nWordsF, h->phiBlockOff, h->sparsePhiTol);
    h->launches++;
    CK(cudaMemcpyAsync(h->hPhiBits2[buf], h->phiBits, sizeof(unsigned int) * h->nWordsF, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->hPhiBlockOff2[buf], h->phiBlockOff, sizeof(int) * ((size_t)h->nPhiBlocks + 1), cudaMemcpyDeviceToHost, s));
}

// phi of this step, sparse form: enqueue the face bitmap (streamD) and its read-back; see k_phi_need_bits.
void sparsePhiBegin(svof_handle* h)
{
    if (!h->phiBits) {
        h->nWordsF = cdiv(h->nF, 32);
        h->nPhiBlocks = cdiv(h->nWordsF, 32);
        h->capPhiPacked = std::max<size_t>(1 << 20, (size_t)h->nF / 4);   // more marked faces than this: the full copy is cheaper
        h->phiBits = dalloc<unsigned int>(h, (size_t)h->nPhiBlocks * 32);
        h->phiBlockOff = dalloc<int>(h, (size_t)h->nPhiBlocks + 1);
        h->phiPacked = dalloc<double>(h, h->capPhiPacked, false);
        for (int i = 0; i < 2; ++i) {
            CK(cudaMallocHost((void**)&h->hPhiBits2[i], sizeof(unsigned int) * (size_t)h->nPhiBlocks * 32));
            CK(cudaMallocHost((void**)&h->hPhiBlockOff2[i], sizeof(int) * ((size_t)h->nPhiBlocks + 1)));
            memset(h->hPhiBits2[i], 0, sizeof(unsigned int) * (size_t)h->nPhiBlocks * 32);
        }
        CK(cudaMallocHost((void**)&h->hPhiPacked, sizeof(double) * h->capPhiPacked));
        h->phiPrevBitsValid = false;
        if (!h->evBits) CK(cudaEventCreateWithFlags(&h->evBits, cudaEventDisableTiming));
        CK(cudaStreamSynchronize(h->stream));   // the zero fills above are ordered on the main stream
    }
    h->pb ^= 1;                    // the bitmap of this call lives in the other host buffer: the previous call's stays readable
    if (h->phiBitsReady) return;   // prefetched (into that buffer) at the end of the previous call
    h->phiPrevBitsValid = false;   // alpha was changed behind the host path: the old bitmap describes nothing the caller holds
    enqueuePhiBits(h, h->streamD, h->pb);
    CK(cudaEventRecord(h->evBits, h->streamD));
}

// ... gather the marked entries of the caller's phi (host threads), copy them up and scatter them (streamD).
// Returns false when too many faces are marked: the caller copies the full field instead.
bool sparsePhiFinish(svof_handle* h, const double* phi, double* zeroStaleIn)
{
    if (!h->phiBitsReady) {
        CK(cudaEventSynchronize(h->evBits));
        hostTick(h, "0a wait: phi face bitmap D2H");
    }
    h->phiBitsReady = false;
    h->d2hBytes += 4LL * h->nWordsF + 4LL * (h->nPhiBlocks + 1);
    HostPool& pool = hostPool(h);
    const int nB = h->nPhiBlocks, nT = pool.size();
    const unsigned int* bits = h->hPhiBits2[h->pb];
    int* off = h->hPhiBlockOff2[h->pb];   // arrives as counts in off[1 + b]: exclusive prefix sum in place
    off[0] = 0;
    long long total = 0;
    for (int b = 0; b < nB; ++b) {
        total += off[b + 1];
        if (total > (long long)h->capPhiPacked) return false;
        off[b + 1] = (int)total;
    }
    // equal shares of MARKED faces per thread (the marked faces cluster where the liquid is): block boundaries from the prefix sums
    h->phiTotal = total;
    std::vector<int>& cut = h->phiCut;
    cut.assign(nT + 1, nB);
    cut[0] = 0;
    for (int t = 1; t < nT; ++t) cut[t] = (int)(std::lower_bound(off, off + nB + 1, (int)(total * t / nT)) - off);
    for (int t = 1; t <= nT; ++t) cut[t] = std::max(cut[t], cut[t - 1]);
    cut[nT] = nB;
    double* packed = h->hPhiPacked;
    const int nWordsF = h->nWordsF;
    const unsigned int* prevBits = (zeroStaleIn && h->phiPrevBitsValid) ? h->hPhiBits2[h->pb ^ 1] : nullptr;
    pool.run(nT, [&](int t) {
        if (prevBits) {   // faces the previous call marked and this one does not: both cells are exactly empty now, their alphaPhi is zero
            const int wa = (int)((long long)nWordsF * t / nT), wb = (int)((long long)nWordsF * (t + 1) / nT);
            for (int w = wa; w < wb; ++w) {
                unsigned int x = prevBits[w] & ~bits[w];
                double* dst = zeroStaleIn + ((size_t)w << 5);
                while (x) {
                    dst[__builtin_ctz(x)] = 0.0;
                    x &= x - 1;
                }
            }
        }
        const int b0 = std::min(cut[t], nB), b1 = std::min(cut[t + 1], nB);
        size_t pos = (size_t)off[b0];
        const int w1 = std::min(32 * b1, nWordsF);
        for (int w = 32 * b0; w < w1; ++w) {
            unsigned int x = bits[w];
            const double* src = phi + ((size_t)w << 5);
            if (x == 0xffffffffu) {   // the marked faces come in long runs (faces of consecutive cells)
                memcpy(packed + pos, src, 32 * sizeof(double));
                pos += 32;
                continue;
            }
            while (x) {
                packed[pos++] = src[__builtin_ctz(x)];
                x &= x - 1;
            }
        }
    });
    hostTick(h, "0c host gather of the marked phi entries");
    if (h->prof) {
        static int once = 0;
        if (!once++) fprintf(stderr, "[svof profile] sparse phi: %lld of %d faces marked, %d host threads\n", total, h->nF, nT);
    }
    CK(cudaMemcpyAsync(h->phiBlockOff, off, sizeof(int) * ((size_t)nB + 1), cudaMemcpyHostToDevice, h->streamD));
    if (total) CK(cudaMemcpyAsync(h->phiPacked, packed, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice, h->streamD));
    k_phi_scatter<<<cdiv(h->nWordsF, 256), 256, 0, h->streamD>>>(h->phiBits, h->phiBlockOff, h->nWordsF, h->phiPacked, h->phi);
    h->launches++;
    h->h2dBytes += 8LL * total + 4LL * (nB + 1);
    return true;
}
// host side of the delta read-back: out[idx[i]] = val[i] (random writes into a field-sized array)
void scatterDeltas(svof_handle* h, const int* idx, const double* val, int cnt, double* out)
{
    if (cnt <= 0) return;
    if (cnt < (1 << 14)) {
        for (int i = 0; i < cnt; ++i) out[idx[i]] = val[i];
        return;
    }
    HostPool& pool = hostPool(h);
    const int nT = pool.size();
    pool.run(nT, [&](int t) {
        const int lo = (int)((long long)cnt * t / nT), hi = (int)((long long)cnt * (t + 1) / nT);
        for (int i = lo; i < hi; ++i) out[idx[i]] = val[i];
    });
}
// host side of the packed alphaPhi read-back: out[f] = packed[k] over the marked faces of this call's bitmap, in face order
void scatterPacked(svof_handle* h, double* out)
{
    HostPool& pool = hostPool(h);
    const int nB = h->nPhiBlocks, nT = pool.size(), nWordsF = h->nWordsF;
    const unsigned int* bits = h->hPhiBits2[h->pb];
    const int* off = h->hPhiBlockOff2[h->pb];
    const double* packed = h->hPhiPacked;
    const std::vector<int>& cut = h->phiCut;
    pool.run(nT, [&](int t) {
        const int b0 = std::min(cut[t], nB), b1 = std::min(cut[t + 1], nB);
        size_t pos = (size_t)off[b0];
        const int w1 = std::min(32 * b1, nWordsF);
        for (int w = 32 * b0; w < w1; ++w) {
            unsigned int x = bits[w];
            double* dst = out + ((size_t)w << 5);
            if (x == 0xffffffffu) {
                memcpy(dst, packed + pos, 32 * sizeof(double));
                pos += 32;
                continue;
            }
            while (x) {
                dst[__builtin_ctz(x)] = packed[pos++];
                x &= x - 1;
            }
        }
    });
}
// device pointer of a caller buffer that is page-locked, device-accessible host memory (cudaMallocHost / svof_host_alloc /
// cudaHostRegister), else nullptr
void* devPtrOfPinned(svof_handle* h, const void* p)
{
    if (!p) return nullptr;
    auto it = h->pinnedCache.find(p);
    if (it != h->pinnedCache.end()) return it->second;
    if (h->pinnedCache.size() > 64) h->pinnedCache.clear();
    void* d = nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess) {
        if (a.type == cudaMemoryTypeHost && a.devicePointer) d = a.devicePointer;
    } else {
        (void)cudaGetLastError();
    }
    h->pinnedCache[p] = d;
    return d;
}

// svof_step_host when every caller buffer is pinned: nothing is staged and the host never waits inside the step.
//   second stream : k_phi_pull (face bitmap of the entries that can matter + their values straight from the caller's phi), Ub
//   main stream   : reconstruct -> U rows next to cut cells pulled from the caller's U -> advect ->
//                   changed alpha cells and the marked alphaPhi faces written straight into the caller's buffers
// One synchronisation at the end (control block: capacity flags and the byte counts).
// An EMPTY cell went out of bounds (overfilled at Courant > 1), so the bounding read phi on faces the sparse upload had
// skipped: rewind to alpha.oldTime (its buffer is intact) and redo the step with the full flux field and full read-backs.
int redoStepWithFullPhi(svof_handle* h, double dt, const double* phi, const double* U, const double* Ub, double* alpha_out,
                        double* alpha_phi_out)
{
    h->cur ^= 1;
    h->cb ^= 1;
    h->bitsValid = false;   // the mixed-cell bitmap on the device belongs to the discarded result
    h->advected = false;
    --h->advectCount;
    h->hostAlphaSynced = h->hostAlphaPhiSynced = nullptr;
    h->phiBitsReady = h->phiPrevBitsValid = h->alphaPhiPrevValid = h->zcPrevValid = false;
    const bool sp = h->sparsePhi;
    h->sparsePhi = false;
    const int rc = svof_step_host(h, dt, phi, U, Ub, alpha_out, alpha_phi_out);
    h->sparsePhi = sp;
    return rc;
}

int stepHostZeroCopy(svof_handle* h, double dt, const double* phi, const double* phiD, const double* U, const double* UD, const double* Ub,
                     double* alpha_out, double* alphaOutD, double* alpha_phi_out, double* alphaPhiOutD)
{
    cudaStream_t st = h->stream, sD = h->streamD;
    h->h2dBytes = h->d2hBytes = 0;
    hostTick(h, nullptr);
    if (!h->zcBits[0]) {
        h->nWordsF = cdiv(h->nF, 32);
        h->zcBits[0] = dalloc<unsigned int>(h, (size_t)h->nWordsF + 1);
        h->zcBits[1] = dalloc<unsigned int>(h, (size_t)h->nWordsF + 1);
        h->zcPrevValid = false;
        CK(cudaStreamSynchronize(st));
    }
    CK(cudaEventRecord(h->evInputs, st));
    CK(cudaStreamWaitEvent(sD, h->evInputs, 0));   // previous step's readers of phi/Ub are done
    const bool pushF = alpha_phi_out && h->hostAlphaPhiSynced == alpha_phi_out && h->zcPrevValid;
    h->zcCur ^= 1;
    unsigned int* bitsCur = h->zcBits[h->zcCur];
    const unsigned int* bitsPrev = h->zcBits[h->zcCur ^ 1];
    CK(cudaMemsetAsync(&h->ctl->nPhiPulled, 0, sizeof(int), sD));
    k_phi_pull<<<cdiv((long long)h->nWordsF * 32, 256), 256, 0, sD>>>(h->md, h->alphaBuf[h->cur], h->sparsePhiTol, phiD, h->phi, bitsCur,
                                                                     h->nWordsF, h->ctl);
    h->launches++;
    if (Ub && h->nBF) {
        CK(cudaMemcpyAsync(h->Ub, Ub, sizeof(double) * 3 * h->nBF, cudaMemcpyHostToDevice, sD));
        h->h2dBytes += 24LL * h->nBF;
    }
    CK(cudaEventRecord(h->evCopy, sD));
    h->havePhi = true;
    h->phiPartial = true;
    if (h->anyInletOutlet && h->nBF) {   // as in the staged path: patch values follow the sign of the new boundary phi
        CK(cudaMemcpyAsync(h->phi + h->nIF, phi + h->nIF, sizeof(double) * h->nBF, cudaMemcpyHostToDevice, st));
        h->h2dBytes += 8LL * h->nBF;
        alphaBC(h);
    }
    // overlapped form: needs the two-stream schedule; isoRDF synchronises with the host inside reconstruct()
    const int ovm = (h->overlap && h->prm.orientation_method != SVOF_ORIENT_ISO_RDF) ? h->zcOverlap : 0;
    const bool ovU = ovm & 1, ovD = ovm & 2, ovP = ovm & 4, ovN = (ovm & 8) && (ovm & 1);
    cudaStream_t sU = h->streamU;
    CK(cudaMemsetAsync(&h->ctl->nDeltaA, 0, 2 * sizeof(int), st));   // nDeltaA, nDeltaF: before anything that may push
    h->zcDenseEarly = ovU || ovD;   // (the front event is recorded for either)
    EventPair& e0 = beginTimed(h, 0);
    doReconstruct(h);
    endTimed(h, e0);
    // U rows the interface-velocity interpolation reads: bitmap only (no list to overflow, nothing to read back)
    if (ovU) {
        // ... marked from the interface-cell list (every interface cell: a superset of the cut ones) as soon as the front
        // kernels are done, and pulled on a third stream while the normals and the plane positioning run
        CK(cudaStreamWaitEvent(sU, h->evFront, 0));
        CK(cudaMemsetAsync(h->uBits, 0, sizeof(unsigned int) * (h->nWords + 1), sU));
        CK(cudaMemsetAsync(&h->ctl->nUCells, 0, sizeof(int), sU));
        k_mark_u_cells<<<sparseGrid(h, 256), 256, 0, sU>>>(h->md, h->mixedCells, nullptr, h->ctl, h->uBits, h->uList, 0);
        k_u_pull<<<cdiv(h->nC, 256), 256, 0, sU>>>(h->uBits, h->nC, UD, h->U);
        h->launches += 2;
        CK(cudaMemsetAsync(h->uBits, 0, sizeof(unsigned int) * (h->nWords + 1), sU));
        CK(cudaEventRecord(h->evU, sU));
        CK(cudaStreamWaitEvent(st, h->evU, 0));
        if (ovN) {   // phi on every face of the needBounding cells, once the bitmap kernel is done with the words
            CK(cudaStreamWaitEvent(sU, h->evCopy, 0));
            k_phi_pull_near<<<sparseGrid(h, 128), 128, 0, sU>>>(h->md, h->near2List, h->near1, h->ctl, phiD, h->phi, bitsCur);
            h->launches++;
            CK(cudaEventRecord(h->evPhiNear, sU));
            CK(cudaStreamWaitEvent(st, h->evPhiNear, 0));
        }
    } else {
        CK(cudaMemsetAsync(h->uBits, 0, sizeof(unsigned int) * (h->nWords + 1), st));
        CK(cudaMemsetAsync(&h->ctl->nUCells, 0, sizeof(int), st));
        LAUNCH(h, k_mark_u_cells, sparseGrid(h, 256), 256, h->md, h->mixedCells, h->cellStatus, h->ctl, h->uBits, h->uList, 0);
        LAUNCH(h, k_u_pull, cdiv(h->nC, 256), 256, h->uBits, h->nC, UD, h->U);
        CK(cudaMemsetAsync(h->uBits, 0, sizeof(unsigned int) * (h->nWords + 1), st));   // leave the bitmap clean for the staged path
    }
    h->haveU = true;
    h->uPartial = true;
    h->inputsAfterNear = true;
    CK(cudaStreamWaitEvent(st, h->evCopy, 0));
    CK(cudaMemsetAsync(&h->ctl->packUnsafe, 0, sizeof(int), st));
    CK(cudaMemsetAsync(&h->ctl->phiUnsafe, 0, sizeof(int), st));
    h->boundPhiBits = bitsCur;   // an out-of-bounds cell with a face outside the bitmap: k_bound_deps reads that entry from the caller's phi
    h->boundPhiHost = phiD;
    h->zcDenseEarly = ovD;
    EventPair& e1 = beginTimed(h, 1);
    doAdvect(h, dt, nullptr, nullptr);
    endTimed(h, e1);
    h->zcDenseEarly = false;
    h->boundPhiBits = nullptr;
    h->boundPhiHost = nullptr;
    // results
    const bool pushA = alpha_out && h->hostAlphaSynced == alpha_out;
    const double* aCur = h->alphaBuf[h->cur];
    const double* aRef = h->alphaBuf[h->cur ^ 1];
    if (ovP && (pushA || pushF)) {
        // final after the streaming kernel: pushed on its stream while the interface chain is still at work ...
        if (ovN) CK(cudaStreamWaitEvent(sD, h->evPhiNear, 0));   // the bitmap words are final
        if (pushA) { k_alpha_push_early<<<h->zcPushCtas * h->sms, 256, 0, sD>>>(aCur, aRef, h->near2, h->nC, alphaOutD, h->ctl); h->launches++; }
        if (pushF) {
            k_alphaphi_push_early<<<h->zcPushCtas * h->sms, 256, 0, sD>>>(h->md, bitsCur, bitsPrev, h->near2, h->alphaPhi, alphaPhiOutD, h->nF, h->ctl);
            h->launches++;
        }
        CK(cudaEventRecord(h->evPush, sD));
        // ... final after k_near_finalize: the near2 cells and the faces they own, from the list
        if (pushA) LAUNCH(h, k_alpha_push_late, sparseGrid(h, 128), 128, h->near2List, h->ctl, aCur, aRef, alphaOutD);
        if (pushF) {
            LAUNCH(h, k_alphaphi_push_late, sparseGrid(h, 128), 128, h->md, h->near2List, h->ctl, bitsCur, bitsPrev, h->alphaPhi, alphaPhiOutD);
            CK(cudaStreamWaitEvent(st, h->evPush, 0));
            LAUNCH(h, k_alphaphi_push_all, 4 * h->sms, 256, h->alphaPhi, alphaPhiOutD, h->nF, h->ctl);
        } else {
            CK(cudaStreamWaitEvent(st, h->evPush, 0));
        }
    } else {
        if (pushA) LAUNCH(h, k_alpha_push, cdiv(h->nC, 256), 256, h->alphaBuf[h->cur], h->alphaBuf[h->cur ^ 1], h->nC, alphaOutD, h->ctl);
        if (pushF) LAUNCH(h, k_alphaphi_push, cdiv(h->nF, 256), 256, bitsCur, bitsPrev, h->alphaPhi, alphaPhiOutD, h->nF, h->ctl);
    }
    if (!pushA && alpha_out) {
        CK(cudaMemcpyAsync(alpha_out, h->alphaBuf[h->cur], sizeof(double) * h->nC, cudaMemcpyDeviceToHost, st));
        h->d2hBytes += 8LL * h->nC;
    }
    if (!pushF && alpha_phi_out) {
        CK(cudaMemcpyAsync(alpha_phi_out, h->alphaPhi, sizeof(double) * h->nF, cudaMemcpyDeviceToHost, st));
        h->d2hBytes += 8LL * h->nF;
    }
    CK(cudaMemcpyAsync(h->hctl, h->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    hostTick(h, "zero-copy step: the one wait");
    h->h2dBytes += 8LL * h->hctl->nPhiPulled + 24LL * std::max(h->hctl->nUCells, 0);
    h->d2hBytes += sizeof(Ctl) + (pushA ? 8LL * h->hctl->nDeltaA : 0) + (pushF ? 8LL * h->hctl->nDeltaF : 0);
    if (alpha_out) h->hostAlphaSynced = alpha_out;
    if (alpha_phi_out) h->hostAlphaPhiSynced = alpha_phi_out;
    h->zcPrevValid = alpha_phi_out != nullptr;
    h->phiBitsReady = h->phiPrevBitsValid = h->alphaPhiPrevValid = false;   // the staged path's mirrors are stale now
    CK(cudaGetLastError());
    if (h->hctl->err) return deviceErr(h);
    if (h->hctl->phiUnsafe) return redoStepWithFullPhi(h, dt, phi, U, Ub, alpha_out, alpha_phi_out);
    return SVOF_OK;
}
}  // namespace

int svof_step_host(svof_handle* h, double dt, const double* phi, const double* U, const double* Ub, double* alpha_out,
                   double* alpha_phi_out)
{
    if (!h || !phi || !U || !(dt > 0)) return SVOF_ERR_INVALID_ARG;
    if (!h->haveAlpha) return fail(h, SVOF_ERR_STATE, "svof_step_host: alpha not set");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    h->zcDenseEarly = false;   // (a failed call may have left it set)
    // the run-time schedule selection belongs to svof_step_device: the host step keeps the schedule as configured
    struct SchedGuard {
        svof_handle* h; int f, d; bool on;
        explicit SchedGuard(svof_handle* h_) : h(h_), f(h_->forkAt), d(h_->denseCtas), on(h_->tuneMode && !h_->schedUser && h_->tuneSlot != 0)
        { if (on) { h->forkAt = h->fork0; h->denseCtas = h->dense0; } }
        ~SchedGuard() { if (on) { h->forkAt = f; h->denseCtas = d; } }
    } schedGuard(h);
    cudaStream_t st = h->stream;
    // decomposed runs: the sparse-phi forms may redo a step on ONE rank (redoStepWithFullPhi), which would unpair the ranks'
    // NCCL ghost refreshes -- full flux field there
    if (h->sparseIO && h->sparsePhi && h->zeroCopy && !h->halo.active) {   // every caller buffer pinned: no staging at all
        void* phiD = devPtrOfPinned(h, phi);
        void* UD = devPtrOfPinned(h, U);
        void* aD = devPtrOfPinned(h, alpha_out);
        void* apD = devPtrOfPinned(h, alpha_phi_out);
        if (phiD && UD && (!alpha_out || aD) && (!alpha_phi_out || apD))
            return stepHostZeroCopy(h, dt, phi, (const double*)phiD, U, (const double*)UD, Ub, alpha_out, (double*)aD, alpha_phi_out,
                                    (double*)apD);
    }
    h->zcPrevValid = false;
    h->h2dBytes = h->d2hBytes = 0;
    hostTick(h, nullptr);
    // phi (the one field that has to cross PCIe in full) goes up on the second stream while reconstruct() and the
    // sparse-U round trip run on the main one
    CK(cudaEventRecord(h->evInputs, st));
    CK(cudaStreamWaitEvent(h->streamD, h->evInputs, 0));  // previous step's readers of phi/Ub are done
    // phi goes up on the second stream while reconstruct() and the sparse-U round trip run on the main one.  Sparse form
    // (default): only the faces whose value can matter (k_phi_need_bits); the device's other entries keep older values,
    // which multiply an exactly zero alpha.  Otherwise the full field (nF doubles) crosses PCIe.
    const bool trySparsePhi = h->sparseIO && h->sparsePhi && !h->halo.active;   // the device's phi starts zero-filled: stale entries are finite
    if (trySparsePhi) sparsePhiBegin(h);
    else {
        CK(cudaMemcpyAsync(h->phi, phi, sizeof(double) * h->nF, cudaMemcpyHostToDevice, h->streamD));
        h->h2dBytes += 8LL * h->nF;
        h->phiPartial = false;
    }
    if (Ub && h->nBF) {
        CK(cudaMemcpyAsync(h->Ub, Ub, sizeof(double) * 3 * h->nBF, cudaMemcpyHostToDevice, h->streamD));
        h->h2dBytes += 24LL * h->nBF;
    }
    if (h->anyInletOutlet && h->nBF) {
        // inletOutlet patch values of alpha.oldTime follow the sign of the NEW phi (what svof_set_phi does for the split
        // calls) and the LS normals read them: the boundary part of phi goes up first, on the main stream
        CK(cudaMemcpyAsync(h->phi + h->nIF, phi + h->nIF, sizeof(double) * h->nBF, cudaMemcpyHostToDevice, st));
        h->h2dBytes += 8LL * h->nBF;
        alphaBC(h);
    }
    EventPair& e0 = beginTimed(h, 0);
    doReconstruct(h);
    endTimed(h, e0);
    bool packedF = false;   // alphaPhi read-back packed by the phi bitmap (set by finishPhi)
    auto finishPhi = [&]() {   // host work that overlaps the reconstruct / U-marking kernels already enqueued
        if (trySparsePhi) {
            // alphaPhi comes back packed by this call's bitmap when the caller's buffer mirrors the previous result
            const bool packedBack = alpha_phi_out && h->hostAlphaPhiSynced == alpha_phi_out && h->phiPrevBitsValid;
            if (sparsePhiFinish(h, phi, packedBack ? alpha_phi_out : nullptr)) { h->phiPartial = true; packedF = packedBack; }
            else {
                CK(cudaMemcpyAsync(h->phi, phi, sizeof(double) * h->nF, cudaMemcpyHostToDevice, h->streamD));
                h->h2dBytes += 8LL * h->nF;
                h->phiPartial = false;
            }
        }
        CK(cudaEventRecord(h->evCopy, h->streamD));
        h->havePhi = true;
    };
    if (!h->sparseIO) finishPhi();
    // U: only the rows the interface-velocity interpolation reads
    bool uDone = false;
    if (h->sparseIO) {
        const int g256 = sparseGrid(h, 256);
        LAUNCH(h, k_clear_u_bits, g256, 256, h->uList, h->ctl, h->uBits, h->capU);
        CK(cudaMemsetAsync(&h->ctl->nUCells, 0, sizeof(int), st));
        LAUNCH(h, k_mark_u_cells, g256, 256, h->md, h->mixedCells, h->cellStatus, h->ctl, h->uBits, h->uList, h->capU);
        // count and a speculative prefix of the list in ONE round trip (sized from the previous step; the rest, if any, follows)
        const int estU = std::min(h->capU, std::max(1 << 14, h->lastNU + h->lastNU / 4));
        CK(cudaMemcpyAsync(&h->hctl->nUCells, &h->ctl->nUCells, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h->hUList, h->uList, sizeof(int) * estU, cudaMemcpyDeviceToHost, st));
        finishPhi();
        CK(cudaStreamSynchronize(st));
        hostTick(h, "1 wait: reconstruct + U-row marking + list D2H");
        const int nU = h->hctl->nUCells;
        h->lastNU = nU;
        if (nU <= h->capU) {
            if (nU > estU) {
                CK(cudaMemcpyAsync(h->hUList + estU, h->uList + estU, sizeof(int) * (nU - estU), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                hostTick(h, "2 U-row list D2H (remainder)");
            }
            if (nU) {
                const int nT = nU > (1 << 15) ? hostPool(h).size() : 1;
                auto gather = [&](int t) {
                    const int lo = (int)((long long)nU * t / nT), hi = (int)((long long)nU * (t + 1) / nT);
                    for (int i = lo; i < hi; ++i) {
                        const double* src = U + 3 * (size_t)h->hUList[i];
                        double* dst = h->hUPacked + 3 * (size_t)i;
                        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
                    }
                };
                if (nT == 1) gather(0); else hostPool(h).run(nT, gather);
                hostTick(h, "3 host gather of U rows");
                CK(cudaMemcpyAsync(h->uPacked, h->hUPacked, sizeof(double) * 3 * nU, cudaMemcpyHostToDevice, st));
                LAUNCH(h, k_scatter_u, cdiv(nU, 256), 256, h->uList, nU, h->uPacked, h->U);
                h->h2dBytes += 24LL * nU;
                h->d2hBytes += 4LL * std::max(nU, estU) + 4;
            }
            uDone = true;
            h->uPartial = true;
        } else {
            CK(cudaMemsetAsync(h->uBits, 0, sizeof(unsigned int) * h->nWords, st));  // list overflowed: bits were left set
        }
    }
    if (!uDone) {
        CK(cudaMemcpyAsync(h->U, U, sizeof(double) * 3 * h->nC, cudaMemcpyHostToDevice, st));
        h->h2dBytes += 24LL * h->nC;
        h->uPartial = false;
    }
    h->haveU = true;
    h->inputsAfterNear = true;  // U landed after the near bitmaps were published
    CK(cudaStreamWaitEvent(st, h->evCopy, 0));  // phi/Ub are on the device
    const bool sparsePhiUsed = trySparsePhi && h->phiPartial;
    CK(cudaMemsetAsync(&h->ctl->phiUnsafe, 0, sizeof(int), st));
    if (sparsePhiUsed) {
        CK(cudaMemsetAsync(&h->ctl->packUnsafe, 0, sizeof(int), st));
        h->boundPhiBits = h->phiBits;   // see k_bound_deps / k_bound_apply
    }
    EventPair& e1 = beginTimed(h, 1);
    doAdvect(h, dt, nullptr, nullptr);
    endTimed(h, e1);
    h->boundPhiBits = nullptr;
    h->boundPhiHost = nullptr;
    // results: deltas against what the caller's buffers already hold, else full copies.  Both delta kernels, the control
    // block (with the two counts) and a speculative prefix of each (index, value) list -- sized from the previous step --
    // are enqueued together and waited for ONCE; a list that turned out longer gets its remainder in a second copy.
    const int capA = h->capDelta / 4, capF = h->capDelta - capA;   // alpha deltas in [0, capA), alphaPhi deltas behind them
    const bool deltaA = alpha_out && h->sparseIO && h->hostAlphaSynced == alpha_out;
    // alphaPhi: packed by this call's phi bitmap when possible (packedF), else deltas against alphaPhiPrev, else the full field
    const bool deltaF = !packedF && alpha_phi_out && h->sparseIO && h->hostAlphaPhiSynced == alpha_phi_out && h->alphaPhiPrevValid;
    int estA = 0, estF = 0;
    if (deltaA || deltaF) CK(cudaMemsetAsync(&h->ctl->nDeltaA, 0, 2 * sizeof(int), st));
    if (deltaA)
        LAUNCH(h, k_delta, cdiv(h->nC, 256), 256, h->alphaBuf[h->cur], (double*)nullptr, h->alphaBuf[h->cur ^ 1], (long long)h->nC,
               &h->ctl->nDeltaA, h->dIdx, h->dVal, capA);
    if (deltaF)
        LAUNCH(h, k_delta, cdiv(h->nF, 256), 256, h->alphaPhi, h->alphaPhiPrev, (const double*)nullptr, (long long)h->nF, &h->ctl->nDeltaF,
               h->dIdx + capA, h->dVal + capA, capF);
    if (packedF) {   // before the device bitmap is overwritten below
        LAUNCH(h, k_phi_pack, cdiv(h->nWordsF, 256), 256, h->phiBits, h->phiBlockOff, h->nWordsF, h->alphaPhi, h->phiPacked);
        if (h->phiTotal) CK(cudaMemcpyAsync(h->hPhiPacked, h->phiPacked, sizeof(double) * (size_t)h->phiTotal, cudaMemcpyDeviceToHost, st));
        h->d2hBytes += 8LL * h->phiTotal;
    }
    // the next call's phi face bitmap (of the alpha just computed) rides along with this call's read-back, into the host
    // buffer this call does not use
    const bool prefetchBits = trySparsePhi && h->phiPartial;
    if (prefetchBits) {
        enqueuePhiBits(h, st, h->pb ^ 1);
    }
    CK(cudaMemcpyAsync(h->hctl, h->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    if (deltaA) {
        estA = std::min(capA, std::max(1 << 14, h->lastCntA + h->lastCntA / 4));
        CK(cudaMemcpyAsync(h->hIdx, h->dIdx, sizeof(int) * estA, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h->hVal, h->dVal, sizeof(double) * estA, cudaMemcpyDeviceToHost, st));
    } else if (alpha_out) {
        CK(cudaMemcpyAsync(alpha_out, h->alphaBuf[h->cur], sizeof(double) * h->nC, cudaMemcpyDeviceToHost, st));
        h->d2hBytes += 8LL * h->nC;
    }
    if (deltaF) {
        estF = std::min(capF, std::max(1 << 14, h->lastCntF + h->lastCntF / 4));
        CK(cudaMemcpyAsync(h->hIdx + capA, h->dIdx + capA, sizeof(int) * estF, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h->hVal + capA, h->dVal + capA, sizeof(double) * estF, cudaMemcpyDeviceToHost, st));
    } else if (alpha_phi_out && !packedF) {
        CK(cudaMemcpyAsync(alpha_phi_out, h->alphaPhi, sizeof(double) * h->nF, cudaMemcpyDeviceToHost, st));
        if (h->sparseIO) CK(cudaMemcpyAsync(h->alphaPhiPrev, h->alphaPhi, sizeof(double) * h->nF, cudaMemcpyDeviceToDevice, st));
        h->d2hBytes += 8LL * h->nF;
        h->alphaPhiPrevValid = h->sparseIO;
    }
    if (packedF) h->alphaPhiPrevValid = false;   // alphaPhiPrev is not maintained by the packed read-back
    CK(cudaStreamSynchronize(st));
    hostTick(h, "4 wait: GPU step + delta kernels + D2H");
    h->d2hBytes += sizeof(Ctl);
    const int cntA = deltaA ? h->hctl->nDeltaA : 0, cntF = deltaF ? h->hctl->nDeltaF : 0;
    bool again = false;
    if (deltaA) {
        h->lastCntA = std::min(cntA, capA);
        if (cntA > capA) {   // too many changes for the list: full copy
            CK(cudaMemcpyAsync(alpha_out, h->alphaBuf[h->cur], sizeof(double) * h->nC, cudaMemcpyDeviceToHost, st));
            h->d2hBytes += 8LL * h->nC;
            again = true;
        } else if (cntA > estA) {
            CK(cudaMemcpyAsync(h->hIdx + estA, h->dIdx + estA, sizeof(int) * (cntA - estA), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(h->hVal + estA, h->dVal + estA, sizeof(double) * (cntA - estA), cudaMemcpyDeviceToHost, st));
            again = true;
        }
        h->d2hBytes += 12LL * std::max(estA, std::min(cntA, capA));
    }
    if (deltaF) {
        h->lastCntF = std::min(cntF, capF);
        if (cntF > capF) {
            CK(cudaMemcpyAsync(alpha_phi_out, h->alphaPhi, sizeof(double) * h->nF, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(h->alphaPhiPrev, h->alphaPhi, sizeof(double) * h->nF, cudaMemcpyDeviceToDevice, st));
            h->d2hBytes += 8LL * h->nF;
            again = true;
        } else if (cntF > estF) {
            CK(cudaMemcpyAsync(h->hIdx + capA + estF, h->dIdx + capA + estF, sizeof(int) * (cntF - estF), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(h->hVal + capA + estF, h->dVal + capA + estF, sizeof(double) * (cntF - estF), cudaMemcpyDeviceToHost, st));
            again = true;
        }
        h->d2hBytes += 12LL * std::max(estF, std::min(cntF, capF));
    }
    if (packedF && h->hctl->packUnsafe) {   // rare: see k_bound_apply
        CK(cudaMemcpyAsync(alpha_phi_out, h->alphaPhi, sizeof(double) * h->nF, cudaMemcpyDeviceToHost, st));
        h->d2hBytes += 8LL * h->nF;
        packedF = false;
        again = true;
    }
    if (again) {
        CK(cudaStreamSynchronize(st));
        hostTick(h, "5 D2H of list remainders / full fields");
    }
    if (deltaA && cntA <= capA) scatterDeltas(h, h->hIdx, h->hVal, cntA, alpha_out);
    if (deltaF && cntF <= capF) scatterDeltas(h, h->hIdx + capA, h->hVal + capA, cntF, alpha_phi_out);
    if (packedF) scatterPacked(h, alpha_phi_out);
    hostTick(h, "6 host scatter");
    if (alpha_out) h->hostAlphaSynced = alpha_out;
    if (alpha_phi_out) h->hostAlphaPhiSynced = alpha_phi_out;
    h->phiBitsReady = prefetchBits;   // valid until anything else changes alpha
    // the bitmap of this call stays in hPhiBits2[pb]: the next call may zero the faces that leave it, provided this call
    // left the caller's alphaPhi buffer complete
    h->phiPrevBitsValid = trySparsePhi && h->phiPartial && alpha_phi_out != nullptr;
    CK(cudaGetLastError());
    if (h->hctl->err) return deviceErr(h);   // a work list or polyhedron cap overflowed: the fields just returned are not to be trusted
    if (sparsePhiUsed && h->hctl->phiUnsafe) return redoStepWithFullPhi(h, dt, phi, U, Ub, alpha_out, alpha_phi_out);
    return SVOF_OK;
    API_END(h)
}

int64_t svof_get_field(svof_handle* h, int which, void* dst, int64_t capacity)
{
    if (!h || !dst) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    const int64_t nC = h->nC, nF = h->nF;
    const void* src = nullptr;
    int64_t n = 0;
    size_t esz = 8;
    switch (which) {
        case SVOF_F_ALPHA: src = h->alphaBuf[h->cur]; n = nC; break;
        case SVOF_F_ALPHA_PHI: src = h->alphaPhi; n = nF; break;
        case SVOF_F_DVF: {
            n = nF;
            if (h->advected) {
                LAUNCH(h, k_materialize_dvf, cdiv(nF, 256), 256, h->md, h->alphaBuf[h->cur ^ 1], h->alphaBBuf[h->cb ^ 1], h->phi, h->lastDt, h->near2,
                       h->dVf, h->scratchF);
            } else {
                CK(cudaMemsetAsync(h->scratchF, 0, sizeof(double) * nF, h->stream));
            }
            src = h->scratchF;
            break;
        }
        case SVOF_F_INTERFACE_N: src = h->iN; n = 3 * nC; break;
        case SVOF_F_INTERFACE_D: src = h->iD; n = nC; break;
        case SVOF_F_INTERFACE_C: src = h->iC; n = 3 * nC; break;
        case SVOF_F_INTERFACE_S: src = h->iS; n = 3 * nC; break;
        case SVOF_F_MIXED_CELLS: fetchCtl(h); src = h->mixedCells; n = h->hctl->nMixed; esz = 4; break;
        case SVOF_F_CELL_STATUS: fetchCtl(h); src = h->cellStatus; n = h->hctl->nMixed; esz = 4; break;
        case SVOF_F_FACE_FLATNESS: src = h->md.flat; n = nF; break;
        case SVOF_F_CF: src = h->md.Cf; n = 3 * nF; break;
        case SVOF_F_SF: src = h->md.Sf; n = 3 * nF; break;
        case SVOF_F_C: src = h->md.C; n = 3 * nC; break;
        case SVOF_F_V: src = h->md.V; n = nC; break;
        case SVOF_F_ALPHA_BOUNDARY: src = h->alphaBBuf[h->cb]; n = h->nBF; break;
        case SVOF_F_UN0: fetchCtl(h); src = h->Un0; n = h->hctl->nMixed; break;
        default: return fail(h, SVOF_ERR_INVALID_ARG, "svof_get_field: unknown field");
    }
    if (capacity < n) return fail(h, SVOF_ERR_INVALID_ARG, "svof_get_field: destination too small");
    if (n) CK(cudaMemcpyAsync(dst, src, (size_t)n * esz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->hctl, h->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->hctl->err) return deviceErr(h);   // where the reference would FatalError: never hand out fields computed past a cap
    return n;
    API_END(h)
}

int svof_get_info(svof_handle* h, int which, double* out)
{
    if (!h || !out) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    // number of sweeps the reference's loop executes = leading sweeps whose loop condition holds
    // (advectionTemplates.C:144-198); derived from the recorded min/max and out-of-bounds counts
    auto nSweeps = [&]() {
        fetchCtl(h);
        const Ctl& c = *h->hctl;
        const bool denseOob = (c.maxDense != 0ull || c.minDense != ~0ull) &&
                              (((dunkey(c.maxDense) - 1.0) > SV_ATOL) || (dunkey(c.minDense) < -SV_ATOL));
        int n = 0;
        while (n < h->sp.nAlphaBounds && (denseOob || c.nearOob[n] > 0)) ++n;
        return n;
    };
    auto sweepVals = [&](bool after, bool wantMin) {
        fetchCtl(h);
        const Ctl& c = *h->hctl;
        if (wantMin) return dunkey(std::min(c.minDense, after ? c.minNearF : c.minNear0));
        return dunkey(std::max(c.maxDense, after ? c.maxNearF : c.maxNear0)) - 1.0;
    };
    switch (which) {
        case SVOF_I_N_MIXED: fetchCtl(h); *out = h->hctl->nMixed; return SVOF_OK;
        case SVOF_I_MIN_ALPHA_BEFORE: *out = sweepVals(false, true); return SVOF_OK;
        case SVOF_I_MAX_ALPHA_M1_BEFORE: *out = sweepVals(false, false); return SVOF_OK;
        case SVOF_I_MIN_ALPHA_AFTER: *out = sweepVals(true, true); return SVOF_OK;
        case SVOF_I_MAX_ALPHA_M1_AFTER: *out = sweepVals(true, false); return SVOF_OK;
        case SVOF_I_N_BOUND_SWEEPS: *out = nSweeps(); return SVOF_OK;
        case SVOF_I_RECONSTRUCTION_TIME: harvestEvents(h, true); *out = h->reconTime; return SVOF_OK;
        case SVOF_I_ADVECTION_TIME: harvestEvents(h, true); *out = h->advTime; return SVOF_OK;
        case SVOF_I_ALPHA_MAPPING_TIME: *out = h->mapTime; return SVOF_OK;
        case SVOF_I_VOLUME: {
            LAUNCH(h, k_volume_partial, 1024, 256, h->alphaBuf[h->cur], h->md.V, h->nC, h->partial);
            CK(cudaMemcpyAsync(h->hpartial, h->partial, 1024 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            // fixed pairwise tree on the host: deterministic
            double buf[1024];
            memcpy(buf, h->hpartial, sizeof(buf));
            for (int o = 512; o > 0; o >>= 1)
                for (int i = 0; i < o; ++i) buf[i] += buf[i + o];
            *out = buf[0];
            return SVOF_OK;
        }
        case SVOF_I_VOLUME_OWNED: {
            if (!h->halo.ownedIdx) return fail(h, SVOF_ERR_STATE, "SVOF_I_VOLUME_OWNED: svof_halo_setup has not been called");
            LAUNCH(h, k_volume_partial_list, 1024, 256, h->halo.ownedIdx, h->halo.nOwned, h->alphaBuf[h->cur], h->md.V, h->partial);
            CK(cudaMemcpyAsync(h->hpartial, h->partial, 1024 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            double buf[1024];
            memcpy(buf, h->hpartial, sizeof(buf));
            for (int o = 512; o > 0; o >>= 1)
                for (int i = 0; i < o; ++i) buf[i] += buf[i + o];
            *out = buf[0];
            return SVOF_OK;
        }
        case SVOF_I_HALO_BYTES: *out = 8.0 * h->halo.cells.nRecv; return SVOF_OK;
        case SVOF_I_RDF_ITERATIONS: *out = (double)h->rdfIterations; return SVOF_OK;
        case SVOF_I_SCHEDULE: {
            const double v = 100.0 * h->forkAt + h->denseCtas;
            const bool measuring = h->tuneMode && !h->schedUser && !h->halo.active && h->overlap && h->tunePhase < 4;
            *out = measuring ? -v : v;
            return SVOF_OK;
        }
        case SVOF_I_GPU_LAUNCHES: *out = (double)h->launches; return SVOF_OK;
        case SVOF_I_FLATNESS_MIN: *out = h->flatMin; return SVOF_OK;
        case SVOF_I_FLATNESS_MAX: *out = h->flatMax; return SVOF_OK;
        case SVOF_I_FLATNESS_AVG: *out = h->flatAvg; return SVOF_OK;
        case SVOF_I_DEVICE_BYTES: *out = (double)h->bytes; return SVOF_OK;
        case SVOF_I_ERROR_FLAGS: fetchCtl(h); *out = h->hctl->err; return SVOF_OK;
        case SVOF_I_DENSE_KERNEL_MS: harvestEvents(h, true); *out = h->denseMs; return SVOF_OK;
        case SVOF_I_DENSE_KERNEL_LAUNCHES: harvestEvents(h, true); *out = (double)h->denseLaunches; return SVOF_OK;
        case SVOF_I_N_NEAR: fetchCtl(h); *out = h->hctl->nNear2; return SVOF_OK;
        case SVOF_I_H2D_BYTES: *out = (double)h->h2dBytes; return SVOF_OK;
        case SVOF_I_D2H_BYTES: *out = (double)h->d2hBytes; return SVOF_OK;
    }
    return fail(h, SVOF_ERR_INVALID_ARG, "svof_get_info: unknown item");
    API_END(h)
}

int svof_device_ptr(svof_handle* h, int which, void** dptr)
{
    if (!h || !dptr) return SVOF_ERR_INVALID_ARG;
    switch (which) {
        case SVOF_F_ALPHA: *dptr = h->alphaBuf[h->cur]; return SVOF_OK;
        case SVOF_F_ALPHA_PHI: *dptr = h->alphaPhi; return SVOF_OK;
        case SVOF_F_CF: *dptr = (void*)h->md.Cf; return SVOF_OK;
        case SVOF_F_SF: *dptr = (void*)h->md.Sf; return SVOF_OK;
        case SVOF_F_C: *dptr = (void*)h->md.C; return SVOF_OK;
        case SVOF_F_V: *dptr = (void*)h->md.V; return SVOF_OK;
        case SVOF_F_INTERFACE_N: *dptr = h->iN; return SVOF_OK;
        case SVOF_F_INTERFACE_D: *dptr = h->iD; return SVOF_OK;
        case SVOF_F_INTERFACE_C: *dptr = h->iC; return SVOF_OK;
        case SVOF_F_INTERFACE_S: *dptr = h->iS; return SVOF_OK;
    }
    return fail(h, SVOF_ERR_INVALID_ARG, "svof_device_ptr: field not exposed");
}

int svof_device_touch(svof_handle* h, int which)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    if (which != SVOF_F_ALPHA) return fail(h, SVOF_ERR_INVALID_ARG, "svof_device_touch: only ALPHA is writable in place");
    API_BEGIN
    CK(cudaSetDevice(h->device));
    alphaBC(h);
    h->haveAlpha = true;
    h->bitsValid = false;
    h->advected = false;
    h->hostAlphaSynced = h->hostAlphaPhiSynced = nullptr; h->phiBitsReady = false;   // the caller's buffers no longer mirror the device
    return SVOF_OK;
    API_END(h)
}

int svof_set_option(svof_handle* h, const char* name, int value)
{
    if (!h || !name) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->streamD));
    CK(cudaStreamSynchronize(h->stream));
    // an explicit schedule option switches the run-time selection of svof_step_device off ("sched_auto 1" switches it back on)
    if (!strcmp(name, "sched_auto")) {
        h->tuneMode = value != 0; h->schedUser = false; h->tunePhase = 0; h->tuneSteps = 0;
        if (h->tuneSlot) tuneUse(h, 0);
        return SVOF_OK;
    }
    if (!strcmp(name, "sched_retune")) { h->retune = true; return SVOF_OK; }
    if (!strcmp(name, "overlap")) { schedByUser(h); h->overlap = value; return SVOF_OK; }
    if (!strcmp(name, "fork")) { schedByUser(h); h->forkAt = value; return SVOF_OK; }
    if (!strcmp(name, "dense_ctas")) { schedByUser(h); h->denseCtas = value; return SVOF_OK; }
    if (!strcmp(name, "dense_threads")) { schedByUser(h); h->denseThreads = value; return SVOF_OK; }
    if (!strcmp(name, "dense_l2")) { schedByUser(h); h->denseL2 = value; return SVOF_OK; }
    if (!strcmp(name, "dense_split")) { schedByUser(h); h->denseSplit = value; return SVOF_OK; }
    if (!strcmp(name, "plic_ctas")) { schedByUser(h); h->plicCtas = value; return SVOF_OK; }
    if (!strcmp(name, "un0_group")) { h->un0Group = value != 0; return SVOF_OK; }
    if (!strcmp(name, "bound_lanes")) { h->boundLanes = value != 0; for (auto& g : h->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; } return SVOF_OK; }
    if (!strcmp(name, "profile")) { h->prof = value != 0; return SVOF_OK; }
    if (!strcmp(name, "dense_v4")) { h->denseV4 = value != 0 && h->dsliced.enabled; for (auto& g : h->graphs) g.sched = -1; return SVOF_OK; }
    if (!strcmp(name, "dense_v3")) { h->denseV3 = value != 0; for (auto& g : h->graphs) g.sched = -1; return SVOF_OK; }
    if (!strcmp(name, "zero_copy")) { h->zeroCopy = value != 0; return SVOF_OK; }
    if (!strcmp(name, "zc_overlap")) { h->zcOverlap = value; return SVOF_OK; }
    if (!strcmp(name, "zc_push_ctas")) { h->zcPushCtas = value > 0 ? value : 2; return SVOF_OK; }
    if (!strcmp(name, "sparse_phi_exp")) {
        h->sparsePhiTol = value > 0 ? pow(10.0, -(double)value) : 0.0;
        h->phiBitsReady = false;
        return SVOF_OK;
    }
    if (!strcmp(name, "sparse_phi")) { h->sparsePhi = value != 0; return SVOF_OK; }
    if (!strcmp(name, "sparse_io")) { h->sparseIO = value != 0; h->hostAlphaSynced = h->hostAlphaPhiSynced = nullptr; h->phiBitsReady = false; return SVOF_OK; }
    return fail(h, SVOF_ERR_INVALID_ARG, "svof_set_option: unknown option");
    API_END(h)
}

int svof_get_stream(svof_handle* h, void** stream)
{
    if (!h || !stream) return SVOF_ERR_INVALID_ARG;
    *stream = (void*)h->stream;
    return SVOF_OK;
}

int svof_synchronize(svof_handle* h)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->streamD));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return checkDeviceErr(h);
    API_END(h)
}

int svof_mark(svof_handle* h, int slot)
{
    if (!h || slot < 0 || slot > 7) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->marks[slot], h->stream));
    return SVOF_OK;
    API_END(h)
}

int svof_elapsed_ms(svof_handle* h, int a, int b, double* ms)
{
    if (!h || !ms || a < 0 || a > 7 || b < 0 || b > 7) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->marks[b]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, h->marks[a], h->marks[b]));
    *ms = f;
    return SVOF_OK;
    API_END(h)
}

int svof_last_step_ms(svof_handle* h, double* r, double* a)
{
    if (!h) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    CK(cudaSetDevice(h->device));
    harvestEvents(h, true);
    if (r) *r = h->lastReconMs;
    if (a) *a = h->lastAdvMs;
    return SVOF_OK;
    API_END(h)
}

int svof_host_alloc(int64_t bytes, void** out)
{
    if (!out || bytes <= 0) return SVOF_ERR_INVALID_ARG;
    if (cudaMallocHost(out, (size_t)bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        g_createError = "cudaMallocHost failed";
        return SVOF_ERR_CUDA;
    }
    return SVOF_OK;
}

int svof_host_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) {
        (void)cudaGetLastError();
        return SVOF_ERR_CUDA;
    }
    return SVOF_OK;
}

// ---- geometry primitives ------------------------------------------------------------------------
#define PRIM_PROLOGUE                      \
    CK(cudaSetDevice(h->device));          \
    std::vector<void*> tmp;                \
    auto up = [&](const void* src, size_t bytes) { \
        void* p = nullptr;                 \
        CK(cudaMalloc(&p, std::max<size_t>(bytes, 8))); \
        tmp.push_back(p);                  \
        if (src) CK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, h->stream)); \
        else CK(cudaMemsetAsync(p, 0, std::max<size_t>(bytes, 8), h->stream)); \
        return p;                          \
    };                                     \
    auto down = [&](void* dstp, const void* srcp, size_t bytes) { CK(cudaMemcpyAsync(dstp, srcp, bytes, cudaMemcpyDeviceToHost, h->stream)); }; \
    int* dErr = (int*)up(nullptr, sizeof(int)); \
    int hErr = 0;
#define PRIM_EPILOGUE                                  \
    down(&hErr, dErr, sizeof(int));                    \
    CK(cudaStreamSynchronize(h->stream));              \
    for (void* p : tmp) cudaFree(p);                   \
    CK(cudaGetLastError());                            \
    if (hErr) return fail(h, SVOF_ERR_CAPACITY, "geometry primitive exceeded a compiled capacity"); \
    return SVOF_OK;

int svof_cut_faces(svof_handle* h, int32_t n_polys, int32_t n_verts, const double* pts, const double* normals, const double* dists,
                   int32_t* status, double* centres, double* areas)
{
    if (!h || n_polys < 0 || n_verts < 3 || !pts || !normals || !dists || !status || !centres || !areas) return SVOF_ERR_INVALID_ARG;
    if (n_verts > CapsPoly::MAXFV) return fail(h, SVOF_ERR_CAPACITY, "svof_cut_faces: more than 16 vertices");
    API_BEGIN
    PRIM_PROLOGUE
    double* dp = (double*)up(pts, sizeof(double) * 3 * (size_t)n_polys * n_verts);
    double* dn = (double*)up(normals, sizeof(double) * 3 * n_polys);
    double* dd = (double*)up(dists, sizeof(double) * n_polys);
    int* ds = (int*)up(nullptr, sizeof(int) * n_polys);
    double* dc = (double*)up(nullptr, sizeof(double) * 3 * n_polys);
    double* da = (double*)up(nullptr, sizeof(double) * 3 * n_polys);
    if (n_polys) { GeoLaunch<CapsPoly>::cutFaces(h->stream, n_polys, n_verts, dp, dn, dd, ds, dc, da, dErr); h->launches++; }
    down(status, ds, sizeof(int) * n_polys);
    down(centres, dc, sizeof(double) * 3 * n_polys);
    down(areas, da, sizeof(double) * 3 * n_polys);
    PRIM_EPILOGUE
    API_END(h)
}

int svof_cut_cells(svof_handle* h, int32_t n, const int32_t* cells, const double* normals, const double* dists, int32_t* status,
                   double* vof, double* sub_volume, double* ic, double* ia)
{
    if (!h || n < 0 || !cells || !normals || !dists || !status || !vof || !sub_volume || !ic || !ia) return SVOF_ERR_INVALID_ARG;
    for (int i = 0; i < n; ++i) if (cells[i] < 0 || cells[i] >= h->nC) return fail(h, SVOF_ERR_INVALID_ARG, "svof_cut_cells: cell out of range");
    API_BEGIN
    PRIM_PROLOGUE
    int* dcell = (int*)up(cells, sizeof(int) * n);
    double* dn = (double*)up(normals, sizeof(double) * 3 * n);
    double* dd = (double*)up(dists, sizeof(double) * n);
    int* ds = (int*)up(nullptr, sizeof(int) * n);
    double* dv = (double*)up(nullptr, sizeof(double) * n);
    double* dsv = (double*)up(nullptr, sizeof(double) * n);
    double* dic = (double*)up(nullptr, sizeof(double) * 3 * n);
    double* dia = (double*)up(nullptr, sizeof(double) * 3 * n);
    if (n) {
        // the non-split capacity variant of this mesh
        const int v = (h->variant >= 3) ? 2 : h->variant;
        switch (v) {
            case 0: GeoLaunch<CapsHex>::cutCells(h->stream, h->md, n, dcell, dn, dd, ds, dv, dsv, dic, dia, dErr); break;
            case 1: GeoLaunch<CapsSmall>::cutCells(h->stream, h->md, n, dcell, dn, dd, ds, dv, dsv, dic, dia, dErr); break;
            default: GeoLaunch<CapsPoly>::cutCells(h->stream, h->md, n, dcell, dn, dd, ds, dv, dsv, dic, dia, dErr); break;
        }
        h->launches++;
    }
    down(status, ds, sizeof(int) * n);
    down(vof, dv, sizeof(double) * n);
    down(sub_volume, dsv, sizeof(double) * n);
    down(ic, dic, sizeof(double) * 3 * n);
    down(ia, dia, sizeof(double) * 3 * n);
    PRIM_EPILOGUE
    API_END(h)
}

int svof_find_signed_distance(svof_handle* h, int32_t n, const int32_t* cells, const double* alphas, const double* normals,
                              int32_t* status, double* dists, double* ic, double* ia)
{
    if (!h || n < 0 || !cells || !alphas || !normals || !status || !dists || !ic || !ia) return SVOF_ERR_INVALID_ARG;
    for (int i = 0; i < n; ++i) if (cells[i] < 0 || cells[i] >= h->nC) return fail(h, SVOF_ERR_INVALID_ARG, "svof_find_signed_distance: cell out of range");
    API_BEGIN
    PRIM_PROLOGUE
    int* dcell = (int*)up(cells, sizeof(int) * n);
    double* dal = (double*)up(alphas, sizeof(double) * n);
    double* dn = (double*)up(normals, sizeof(double) * 3 * n);
    int* ds = (int*)up(nullptr, sizeof(int) * n);
    double* dd = (double*)up(nullptr, sizeof(double) * n);
    double* dic = (double*)up(nullptr, sizeof(double) * 3 * n);
    double* dia = (double*)up(nullptr, sizeof(double) * 3 * n);
    if (n) GEO(h, findDistance, h->stream, h->md, n, dcell, dal, dn, h->sp.split, ds, dd, dic, dia, dErr);
    down(status, ds, sizeof(int) * n);
    down(dists, dd, sizeof(double) * n);
    down(ic, dic, sizeof(double) * 3 * n);
    down(ia, dia, sizeof(double) * 3 * n);
    PRIM_EPILOGUE
    API_END(h)
}

int svof_face_fluxes(svof_handle* h, int32_t n, const int32_t* faces, const double* normals, const double* dists, const double* Un0,
                     double dt, const double* phi, double* dVf)
{
    if (!h || n < 0 || !faces || !normals || !dists || !Un0 || !phi || !dVf) return SVOF_ERR_INVALID_ARG;
    for (int i = 0; i < n; ++i) if (faces[i] < 0 || faces[i] >= h->nF) return fail(h, SVOF_ERR_INVALID_ARG, "svof_face_fluxes: face out of range");
    API_BEGIN
    PRIM_PROLOGUE
    int* df = (int*)up(faces, sizeof(int) * n);
    double* dn = (double*)up(normals, sizeof(double) * 3 * n);
    double* dd = (double*)up(dists, sizeof(double) * n);
    double* du = (double*)up(Un0, sizeof(double) * n);
    double* dph = (double*)up(phi, sizeof(double) * n);
    double* dout = (double*)up(nullptr, sizeof(double) * n);
    if (n) GEO(h, faceFluxes, h->stream, h->md, n, df, dn, dd, du, dt, dph, dout, dErr);
    down(dVf, dout, sizeof(double) * n);
    PRIM_EPILOGUE
    API_END(h)
}

int svof_plic_surface(svof_handle* h, int64_t cap_points, int64_t cap_faces, double* points, int32_t* face_offsets, int32_t* cells,
                      int64_t* n_points, int64_t* n_faces)
{
    if (!h || !n_points || !n_faces) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    PRIM_PROLOGUE
    (void)down;
    fetchCtl(h);
    const int nM = h->hctl->nMixed;
    int maxEp = 0;
    switch (h->variant) {
        case 0: maxEp = GeoLaunch<CapsHex>::maxPolyPoints(); break;
        case 1: maxEp = GeoLaunch<CapsSmall>::maxPolyPoints(); break;
        case 2: maxEp = GeoLaunch<CapsPoly>::maxPolyPoints(); break;
        case 4: maxEp = GeoLaunch<CapsHexSplit>::maxPolyPoints(); break;
        default: maxEp = GeoLaunch<CapsSplit>::maxPolyPoints(); break;
    }
    double* dPts = (double*)up(nullptr, sizeof(double) * 3 * (size_t)std::max(nM, 1) * maxEp);
    int* dCnt = (int*)up(nullptr, sizeof(int) * (size_t)std::max(nM, 1));
    std::vector<int> cnt(nM), mixed(nM);
    std::vector<double> pts;
    if (nM) {
        GEO(h, plicPolygons, h->stream, sparseGrid(h, 128), h->md, h->mixedCells, h->ctl, h->iN, h->iD, dPts, dCnt);
        CK(cudaMemcpyAsync(cnt.data(), dCnt, sizeof(int) * nM, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(mixed.data(), h->mixedCells, sizeof(int) * nM, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    int64_t nP = 0, nFc = 0;
    for (int i = 0; i < nM; ++i)
        if (cnt[i] > 0) { nP += cnt[i]; nFc++; }
    *n_points = nP;
    *n_faces = nFc;
    int rc = SVOF_OK;
    if (points) {
        if (cap_points < nP || cap_faces < nFc || !face_offsets || !cells) {
            rc = SVOF_ERR_CAPACITY;
        } else if (nM) {
            pts.resize((size_t)3 * nM * maxEp);
            CK(cudaMemcpyAsync(pts.data(), dPts, sizeof(double) * pts.size(), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            int64_t p = 0, f = 0;
            face_offsets[0] = 0;
            for (int i = 0; i < nM; ++i) {
                if (cnt[i] <= 0) continue;
                std::memcpy(points + 3 * p, pts.data() + (size_t)3 * i * maxEp, sizeof(double) * 3 * cnt[i]);
                p += cnt[i];
                cells[f] = mixed[i];
                face_offsets[++f] = (int32_t)p;
            }
        } else {
            face_offsets[0] = 0;
        }
    }
    fetchCtl(h);
    hErr = h->hctl->err;
    for (void* q : tmp) cudaFree(q);
    CK(cudaGetLastError());
    if (rc != SVOF_OK) return fail(h, rc, "svof_plic_surface: output arrays too small");
    if (hErr) return fail(h, SVOF_ERR_CAPACITY, "svof_plic_surface: a cell exceeded a compiled capacity");
    return SVOF_OK;
    API_END(h)
}


// reconstruction::subCellFaces() (reconstruction.C:838-891).  The polygons come from k_subcell_faces; the point merge and
// the orientation fix of cutCell::updateSubCellPointsandFaces (cutCell.C:239-290) are list surgery on a few thousand small
// polygons and run here on the host (post-processing for the reconstructedSubcellFaces sampler, not part of the step).
int svof_subcell_faces(svof_handle* h, int64_t cap_points, int64_t cap_faces, int64_t cap_face_points, double* points,
                       int32_t* face_offsets, int32_t* face_points, int32_t* face_cell, int64_t* n_points, int64_t* n_faces,
                       int64_t* n_face_points)
{
    if (!h || !n_points || !n_faces || !n_face_points) return SVOF_ERR_INVALID_ARG;
    API_BEGIN
    PRIM_PROLOGUE
    (void)down;
    (void)dErr;
    fetchCtl(h);
    const int nM = h->hctl->nMixed;
    const int maxFaces = h->maxCF + 1;
    const int maxPts = h->maxCF * 2 * h->md.maxFV + 2 * h->maxCF + 8;
    double* dPts = (double*)up(nullptr, sizeof(double) * 3 * (size_t)std::max(nM, 1) * maxPts);
    int* dFs = (int*)up(nullptr, sizeof(int) * (size_t)std::max(nM, 1) * maxFaces);
    int* dNf = (int*)up(nullptr, sizeof(int) * (size_t)std::max(nM, 1));
    double* dCen = (double*)up(nullptr, sizeof(double) * 3 * (size_t)std::max(nM, 1));
    std::vector<double> pts((size_t)3 * nM * maxPts), cen((size_t)3 * nM);
    std::vector<int> fs((size_t)nM * maxFaces), nf(nM), mixed(nM);
    if (nM) {
        const int v = (h->variant >= 3) ? 2 : h->variant;   // evaluated without splitWarpedFace (reconstruction.C:856)
        const int grid = sparseGrid(h, 128);
        switch (v) {
            case 0: GeoLaunch<CapsHex>::subCellFaces(h->stream, grid, h->md, h->mixedCells, h->ctl, h->iN, h->iD, maxFaces, maxPts, dPts, dFs, dNf, dCen); break;
            case 1: GeoLaunch<CapsSmall>::subCellFaces(h->stream, grid, h->md, h->mixedCells, h->ctl, h->iN, h->iD, maxFaces, maxPts, dPts, dFs, dNf, dCen); break;
            default: GeoLaunch<CapsPoly>::subCellFaces(h->stream, grid, h->md, h->mixedCells, h->ctl, h->iN, h->iD, maxFaces, maxPts, dPts, dFs, dNf, dCen); break;
        }
        h->launches++;
        CK(cudaMemcpyAsync(pts.data(), dPts, sizeof(double) * pts.size(), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(cen.data(), dCen, sizeof(double) * cen.size(), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(fs.data(), dFs, sizeof(int) * fs.size(), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(nf.data(), dNf, sizeof(int) * nM, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(mixed.data(), h->mixedCells, sizeof(int) * nM, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    for (void* q : tmp) cudaFree(q);
    fetchCtl(h);
    if (h->hctl->err) return deviceErr(h);
    auto P = [&](const std::vector<d3>& v, int i) -> const d3& { return v[(size_t)i]; };
    // face::centre / face::areaNormal (OF, recalled) on a host polygon, as svof_geom.cuh does on the device
    auto polyCentre = [&](const std::vector<d3>& p) {
        const int n = (int)p.size();
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
    };
    auto polyArea = [&](const std::vector<d3>& p) {
        const int n = (int)p.size();
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
    };
    (void)P;
    std::vector<double> outPts;
    std::vector<int32_t> off(1, 0), fpts, fcell;
    for (int i = 0; i < nM; ++i) {
        if (nf[i] <= 0) continue;
        const double* cp = pts.data() + (size_t)3 * i * maxPts;
        const d3 centre = mk3(cen[3 * (size_t)i], cen[3 * (size_t)i + 1], cen[3 * (size_t)i + 2]);
        // merge duplicate points (first occurrence kept, 10*SMALL), as the oracle does
        std::vector<d3> uniq;
        std::vector<int> toUnique;
        int np = 0;
        for (int k = 0; k < nf[i]; ++k) np += fs[(size_t)i * maxFaces + k];
        toUnique.resize(np);
        for (int q = 0; q < np; ++q) {
            const d3 x = mk3(cp[3 * q], cp[3 * q + 1], cp[3 * q + 2]);
            int found = -1;
            for (size_t j = 0; j < uniq.size() && found < 0; ++j)
                if (mag(x - uniq[j]) <= 10.0 * SV_SMALL) found = (int)j;
            if (found < 0) {
                found = (int)uniq.size();
                uniq.push_back(x);
            }
            toUnique[q] = found;
        }
        const int32_t base = (int32_t)(outPts.size() / 3);
        int q0 = 0;
        for (int k = 0; k < nf[i]; ++k) {
            const int cnt = fs[(size_t)i * maxFaces + k];
            std::vector<int> f(cnt);
            std::vector<d3> fp(cnt);
            for (int q = 0; q < cnt; ++q) {
                f[q] = toUnique[q0 + q];
                fp[q] = uniq[f[q]];
            }
            q0 += cnt;
            if (dot(polyCentre(fp) - centre, polyArea(fp)) < 0.0) std::reverse(f.begin() + 1, f.end());  // face::reverseFace keeps vertex 0
            for (int v : f) fpts.push_back(base + v);
            off.push_back((int32_t)fpts.size());
            fcell.push_back(mixed[i]);
        }
        for (const d3& x : uniq) {
            outPts.push_back(x.x);
            outPts.push_back(x.y);
            outPts.push_back(x.z);
        }
    }
    *n_points = (int64_t)(outPts.size() / 3);
    *n_faces = (int64_t)fcell.size();
    *n_face_points = (int64_t)fpts.size();
    if (!points) return SVOF_OK;
    if (cap_points < *n_points || cap_faces < *n_faces || cap_face_points < *n_face_points || !face_offsets || !face_points || !face_cell)
        return fail(h, SVOF_ERR_CAPACITY, "svof_subcell_faces: output arrays too small");
    std::copy(outPts.begin(), outPts.end(), points);
    std::copy(off.begin(), off.end(), face_offsets);
    std::copy(fpts.begin(), fpts.end(), face_points);
    std::copy(fcell.begin(), fcell.end(), face_cell);
    return SVOF_OK;
    API_END(h)
}

}  // extern "C"
