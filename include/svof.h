/*
 * svof.h -- C ABI of the B200-native SimPLIC volume-fraction transport step.
 *
 * This is the drop-in boundary for the one hot path of daidezhi/geometricVofExt:
 *     geometricVofExt::SimPLIC::solveVofEqu::reconstruct() + advect(Sp, Su)
 * Everything crossing it is a plain pointer + size (SoA, int32 labels, IEEE
 * double scalars: the reference's `arch "LSB;label=32;scalar=64"`).  There are
 * no C++/torch types in any signature, so it can be bound from C++ (the
 * OpenFOAM adapter in INTEGRATION.md), ctypes, cgo or JNI alike.
 *
 * Each entry point cites the reference interface it replaces (paths relative
 * to the reference tree).  Every function returns 0 on success or a negative
 * svof_status; the library never calls exit()/abort() (the reference would
 * FatalError-abort, e.g. reconstruction.C:610-625, advectionTemplates.C:58-63;
 * the adapter turns a non-zero return into FatalErrorInFunction).
 *
 * Two shared objects implement this header with identical semantics:
 *   geometricvofext_b200/lib/libsvof_b200.so   -- the product: CUDA, sm_100a
 *   oracle/_build/libsvof_oracle.so            -- test infrastructure only: a
 *        single-threaded CPU restatement of the reference algorithm
 */
#ifndef SVOF_H
#define SVOF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVOF_ABI_VERSION 2

typedef struct svof_handle svof_handle;

typedef enum {
    SVOF_OK = 0,
    SVOF_ERR_INVALID_ARG = -1,   /* null pointer, bad size, unknown key      */
    SVOF_ERR_BAD_MESH = -2,      /* inconsistent connectivity / patch table  */
    SVOF_ERR_BAD_CONFIG = -3,    /* e.g. invalid orientationMethod           */
    SVOF_ERR_CUDA = -4,          /* CUDA runtime error (text in last_error)  */
    SVOF_ERR_COMM = -5,          /* halo exchange / peer mapping error       */
    SVOF_ERR_CAPACITY = -6,      /* polygon/polyhedron exceeds a device cap  */
    SVOF_ERR_STATE = -7,         /* call order (advect before fields set...) */
    SVOF_ERR_UNSUPPORTED = -8
} svof_status;

/* polyPatch classes that matter to the path (advectionTemplates.C:66-70,
 * advection.C:403,425).  Everything that is neither empty nor processor is
 * GENERIC (wall, patch, ...). */
typedef enum {
    SVOF_PATCH_GENERIC = 0,
    SVOF_PATCH_EMPTY = 1,
    SVOF_PATCH_PROCESSOR = 2
} svof_patch_kind;

/* alpha boundary conditions used by the reference's cases:
 * zeroGradient (tutorials/test/plicVofAdvectionFoam/0.orig/alpha.water),
 * inletOutlet (tutorials/solvers/interPlicFoam/damBreakWithObstacle/0.orig/alpha.water),
 * fixedValue. */
typedef enum {
    SVOF_BC_ZERO_GRADIENT = 0,
    SVOF_BC_FIXED_VALUE = 1,
    SVOF_BC_INLET_OUTLET = 2
} svof_alpha_bc;

typedef struct {
    int32_t start;        /* first face (global face index, >= n_internal_faces) */
    int32_t size;         /* number of faces                                      */
    int32_t kind;         /* svof_patch_kind                                      */
    int32_t nbr_rank;     /* processor patches: neighbour rank, else -1           */
    int32_t alpha_bc;     /* svof_alpha_bc (GENERIC patches)                      */
    int32_t reserved;
    double alpha_value;   /* fixedValue value / inletOutlet inletValue            */
} svof_patch;

/* polyMesh in OpenFOAM's own layout (constant/polyMesh/{points,faces,owner,
 * neighbour,boundary}): faces [0,n_internal_faces) are internal, boundary
 * faces follow patch by patch. */
typedef struct {
    int32_t n_points;
    int32_t n_faces;
    int32_t n_internal_faces;
    int32_t n_cells;
    int32_t n_patches;
    int32_t reserved;
    const double* points;        /* [3*n_points] xyz interleaved              */
    const int32_t* face_offsets; /* [n_faces+1] CSR into face_points          */
    const int32_t* face_points;  /* point labels, owner-outward orientation   */
    const int32_t* owner;        /* [n_faces]                                 */
    const int32_t* neighbour;    /* [n_internal_faces]                        */
    const svof_patch* patches;   /* [n_patches], ascending start              */
    /* Optional host geometry (fvMesh::Cf/Sf/C/V).  NULL => derived with
     * OpenFOAM's primitiveMeshTools formulas. */
    const double* Cf;            /* [3*n_faces] or NULL */
    const double* Sf;            /* [3*n_faces] or NULL */
    const double* C;             /* [3*n_cells] or NULL */
    const double* V;             /* [n_cells]   or NULL */
} svof_mesh;

/* orientationMethod (reconstruction.C:60-68) */
typedef enum {
    SVOF_ORIENT_ALPHA_GRAD = 0,     /* "alphaGrad"                       */
    SVOF_ORIENT_ISO_ALPHA_GRAD = 1, /* "isoAlphaGrad" | "LS"  (default)  */
    SVOF_ORIENT_ISO_RDF = 2         /* "isoRDF" | "RDF"                  */
} svof_orientation;

/* The fvSolution solvers."alpha.*" keys the path reads
 * (reconstruction.C:502-516, advection.C:455-457) plus the two isoAdvector
 * names the north star lists (surfCellTol -> alias of mixedCellTol,
 * isoFaceTol -> accepted and ignored: SimPLIC's plane position is analytic). */
typedef struct {
    double mixed_cell_tol;      /* mixedCellTol   1e-8   */
    double snap_tol;            /* snapTol        0      */
    double iso_face_tol;        /* isoFaceTol     (ignored, recorded)      */
    double rdf_tol;             /* tol            1e-6   (isoRDF only)     */
    double rdf_rel_tol;         /* relTol         0.1    (isoRDF only)     */
    int32_t n_alpha_bounds;     /* nAlphaBounds   10     */
    int32_t clip;               /* clip           true   */
    int32_t orientation_method; /* svof_orientation, default ISO_ALPHA_GRAD */
    int32_t split_warped_face;  /* splitWarpedFace false */
    int32_t map_alpha_field;    /* mapAlphaField  false  */
    int32_t write_plic_fields;  /* writePlicFields false */
    int32_t rdf_iterations;     /* iterations     5      (isoRDF only)     */
    int32_t mixed_cell_tol_set; /* internal: explicit mixedCellTol wins over surfCellTol */
    int32_t alpha_grad_scheme;  /* orientationMethod alphaGrad only: the caller's fvSchemes gradSchemes entry for
                                 * grad(alpha1) (reconstruction.C:78) -- 0 "Gauss linear" (default; every solver
                                 * tutorial), 1 "Gauss pointLinear" (tutorials/test/plicVofOrientationFoam/NAG/system/
                                 * fvSchemes:35).  svof_params_set key "gradSchemes" (or "grad(alpha1)"). */
} svof_params;

/* One handle <-> one rank <-> one GPU (one MPI rank of the reference).
 * world_size > 1: the library opens its own NCCL communicator from nccl_unique_id
 * (128 bytes from svof_comm_unique_id on rank 0, broadcast by whatever transport the
 * host has: MPI_Bcast inside OpenFOAM, torch.distributed in bench.py); svof_create is
 * then a collective call.  See "decomposed runs" below. */
typedef struct {
    int32_t rank;
    int32_t world_size;
    int32_t device;   /* CUDA device ordinal, -1 => rank % deviceCount */
    int32_t reserved;
    const void* nccl_unique_id; /* [128 bytes] or NULL (world_size == 1) */
} svof_comm;

/* ---- configuration -------------------------------------------------------- */

/* Defaults of reconstruction.C:502-516 / advection.C:455-457. */
int svof_params_default(svof_params* p);

/* Set one fvSolution key from its dictionary text ("nAlphaBounds","3"),
 * ("clip","false"), ("orientationMethod","LS"), ("surfCellTol","1e-8")...
 * Unknown keys return SVOF_ERR_INVALID_ARG; keys of the same dictionary that
 * belong to the caller (nAlphaSubCycles, cAlpha, period, reverseTime) are
 * accepted and ignored. */
int svof_params_set(svof_params* p, const char* key, const char* value);

/* ---- life cycle ------------------------------------------------------------ */

/* Replaces the solveVofEqu constructor (solveVofEqu.C:55-91), i.e. the
 * reconstruction ctor (reconstruction.C:478-629, incl. updateFaceFlatness
 * :408-473) and the advection ctor (advection.C:437-522). */
int svof_create(const svof_mesh* mesh, const svof_params* params,
                const svof_comm* comm, svof_handle** out);
int svof_destroy(svof_handle* h);

/* Text of the last error on this handle (or of the last failed svof_create
 * when h == NULL).  Never NULL. */
const char* svof_last_error(const svof_handle* h);

/* ---- fields in (host pointers) ------------------------------------------- */

/* alpha1 internal field [n_cells]; boundary values are evaluated from the patch
 * BCs (volScalarField::correctBoundaryConditions).  Also resets alpha.oldTime. */
int svof_set_alpha(svof_handle* h, const double* alpha);
/* phi: face volume flux [n_faces] (internal then boundary faces, patch order;
 * surfaceScalarField primitive + boundary fields flattened). */
int svof_set_phi(svof_handle* h, const double* phi);
/* U: cell values [3*n_cells] and boundary-face values [3*(n_faces-n_internal)]
 * (volVectorField internal + boundary field after correctBoundaryConditions),
 * consumed by the interface-velocity interpolation (advection.C:91,126). */
int svof_set_U(svof_handle* h, const double* U, const double* Ub);

/* ---- the step --------------------------------------------------------------- */

/* solveVofEqu::reconstruct()  (solveVofEqu.C:96-99 -> reconstruction.C:680-722) */
int svof_reconstruct(svof_handle* h);

/* solveVofEqu::advect(Sp,Su)  (solveVofEquTemplates.C:35-43 ->
 * advectionTemplates.C:352-418).  Sp/Su: [n_cells] or NULL for zeroField
 * (interPlicFoam/alphaSuSp.H:1-2).  alpha.oldTime() is the alpha held at entry. */
int svof_advect(svof_handle* h, double dt, const double* Sp, const double* Su);

/* reconstruct() + advect(dt, zeroField, zeroField) on the fields already on the device, enqueued as ONE CUDA-graph
 * launch in the steady state (same results as the two calls; the per-phase timers are not updated by graph
 * launches).  Inputs are refreshed between calls with svof_set_phi/_U or their _device variants.  CUDA library only. */
int svof_step_device(svof_handle* h, double dt);

/* Host-pointer convenience = set_phi + set_U + reconstruct + advect + read back
 * alpha (and alphaPhi if non-NULL): the end-to-end call bench.py times.  On return the caller's buffers
 * hold the complete new fields (see the "sparse_io" / "sparse_phi" / "zero_copy" options for how few bytes that takes).
 * When phi, U, alpha_out and alpha_phi_out are page-locked, device-accessible host memory (svof_host_alloc, cudaMallocHost,
 * cudaHostRegister) nothing is staged: the kernels read the entries of phi / rows of U they need and write the changed alpha
 * cells and alphaPhi faces straight into the caller's buffers, and the call waits for the device once.  Pageable buffers
 * take the staged path (face bitmap, host gather/scatter on a small thread pool, three round trips).  Either way the result
 * is bitwise that of full-field copies; if an EMPTY cell goes out of bounds (possible only at Courant > 1) the library
 * notices on the device and redoes the step with the full flux field. */
int svof_step_host(svof_handle* h, double dt, const double* phi, const double* U,
                   const double* Ub, double* alpha_out, double* alpha_phi_out);

/* ---- fields out -------------------------------------------------------------- */

typedef enum {
    SVOF_F_ALPHA = 0,        /* f64 [n_cells]                                     */
    SVOF_F_ALPHA_PHI = 1,    /* f64 [n_faces]   advection::alphaPhi()             */
    SVOF_F_DVF = 2,          /* f64 [n_faces]   dVf_ after the step               */
    SVOF_F_INTERFACE_N = 3,  /* f64 [3*n_cells] reconstruction::interfaceN()      */
    SVOF_F_INTERFACE_D = 4,  /* f64 [n_cells]                                     */
    SVOF_F_INTERFACE_C = 5,  /* f64 [3*n_cells]                                   */
    SVOF_F_INTERFACE_S = 6,  /* f64 [3*n_cells]                                   */
    SVOF_F_MIXED_CELLS = 7,  /* i32 [n_mixed]   reconstruction::mixedCells()      */
    SVOF_F_CELL_STATUS = 8,  /* i32 [n_mixed]   reconstruction::cellStatus()      */
    SVOF_F_FACE_FLATNESS = 9,/* f64 [n_faces]   reconstruction::faceFlatness()    */
    SVOF_F_CF = 10,          /* f64 [3*n_faces] mesh geometry as used             */
    SVOF_F_SF = 11,          /* f64 [3*n_faces]                                   */
    SVOF_F_C = 12,           /* f64 [3*n_cells]                                   */
    SVOF_F_V = 13,           /* f64 [n_cells]                                     */
    SVOF_F_ALPHA_BOUNDARY = 14, /* f64 [n_faces-n_internal] alpha patch values    */
    SVOF_F_UN0 = 15,         /* f64 [n_mixed] interface normal speed (advection.C:126) */
    SVOF_F_COUNT_
} svof_field;

/* Copy a field to host memory.  capacity is in ELEMENTS of the field's type;
 * returns the number of elements written (>= 0) or a negative status. */
int64_t svof_get_field(svof_handle* h, int which, void* dst, int64_t capacity);

typedef enum {
    SVOF_I_N_MIXED = 0,          /* mixedCells().size() of the last reconstruct   */
    SVOF_I_MIN_ALPHA_BEFORE = 1, /* "Before conservative bounding" log values     */
    SVOF_I_MAX_ALPHA_M1_BEFORE = 2,
    SVOF_I_MIN_ALPHA_AFTER = 3,  /* "After  conservative bounding"                */
    SVOF_I_MAX_ALPHA_M1_AFTER = 4,
    SVOF_I_N_BOUND_SWEEPS = 5,   /* sweeps executed by limitFlux                  */
    SVOF_I_RECONSTRUCTION_TIME = 6, /* s, cumulative (reconstruction.C:721)       */
    SVOF_I_ADVECTION_TIME = 7,      /* s, cumulative (advectionTemplates.C:415)   */
    SVOF_I_ALPHA_MAPPING_TIME = 8,
    SVOF_I_VOLUME = 9,           /* sum(alpha*V) (plicVof.H:44-46)                */
    SVOF_I_GPU_LAUNCHES = 10,    /* kernels launched so far (0 for the oracle)    */
    SVOF_I_FLATNESS_MIN = 11, SVOF_I_FLATNESS_MAX = 12, SVOF_I_FLATNESS_AVG = 13,
    SVOF_I_DEVICE_BYTES = 14,
    SVOF_I_ERROR_FLAGS = 15,     /* device-side capacity flags, 0 = clean         */
    SVOF_I_DENSE_KERNEL_MS = 16, /* cumulative device ms of the streaming kernel  */
    SVOF_I_DENSE_KERNEL_LAUNCHES = 17,
    SVOF_I_N_NEAR = 18,          /* |mixed U 2 face-neighbour layers| (sparse set) */
    SVOF_I_H2D_BYTES = 19,       /* bytes the last svof_step_host copied host->device */
    SVOF_I_D2H_BYTES = 20,       /* ... and device->host                          */
    SVOF_I_VOLUME_OWNED = 21,    /* sum(alpha*V) over the cells this rank owns (== VOLUME without ghosts) */
    SVOF_I_HALO_BYTES = 22,      /* bytes this rank receives per ghost refresh    */
    SVOF_I_RDF_ITERATIONS = 23,  /* isoRDF: iterations the last reconstruct ran (reconstruction.C:228-399) */
    SVOF_I_SCHEDULE = 24,        /* svof_step_device: schedule in use, 100*fork + resident streaming CTAs per SM (0 = uncapped);
                                  * negative while the run-time selection ("sched_auto") is still measuring */
    SVOF_I_COUNT_
} svof_info;

int svof_get_info(svof_handle* h, int which, double* out);

/* Device-resident access for callers that already live on the GPU (the
 * benchmark's roofline leg).  Returns a CUDA device pointer valid until
 * svof_destroy.  The oracle returns SVOF_ERR_UNSUPPORTED. */
int svof_device_ptr(svof_handle* h, int which, void** dptr);
/* Tell the handle that a device-resident input (ALPHA, via svof_device_ptr) was
 * overwritten in place by the caller. */
int svof_device_touch(svof_handle* h, int which);
/* alpha[idx[i]] = vals[i] for i < n, idx/vals DEVICE arrays, enqueued on the handle's stream: how a decomposed run
 * refreshes its halo cells after the swap (the mixed-cell bitmap and the patch values follow; no dense pass).
 * CUDA library only. */
int svof_scatter_alpha_device(svof_handle* h, const int32_t* d_idx, const double* d_vals, int64_t n);
/* phi/U staged on device already: same as svof_set_phi/svof_set_U but
 * device-to-device (CUDA library only). */
int svof_set_phi_device(svof_handle* h, const void* dphi);
int svof_set_U_device(svof_handle* h, const void* dU, const void* dUb);
/* Runtime switches: "overlap" (default 1: the streaming kernel runs on a second stream beside the interface
 * kernels; 0: one stream -- bitwise the same result either way), "fork", "plic_ctas", "dense_ctas" (tuning of that
 * schedule), "profile" (0/1: CUDA events around every launch, printed at destroy),
 * "sparse_io" (0/1, default 1: svof_step_host uploads only the rows of U the interface-velocity interpolation
 * reads and reads alpha/alphaPhi back as (index,value) deltas against what the SAME caller buffers received
 * from the previous svof_step_host; a caller that modifies those buffers in between must call svof_set_alpha
 * or pass 0),
 * "sparse_phi" (0/1, default 1, with sparse_io: svof_step_host copies up only the phi entries the step can depend
 * on -- every boundary face and every internal face with alpha != 0 in at least one of its two cells; on the other
 * faces phi multiplies an exactly zero alpha, so the device keeps whatever it held.  The caller's phi buffer is read
 * afresh in every call.  After such a call the device's phi is not a full field: svof_advect / svof_step_device
 * return SVOF_ERR_STATE until svof_set_phi / svof_set_phi_device),
 * "zero_copy" (0/1, default 1: the pinned-buffer path of svof_step_host described there),
 * "sparse_phi_exp" (e > 0: for that upload, cells with |alpha| <= 10^-e count as empty.  With snapTol 0 the support of
 * alpha grows one cell layer per step downstream -- round-off-sized values carried by the upwind flux -- until the sparse
 * upload degenerates into the full one; the threshold keeps it tight at the price of an O(10^-e) difference from the
 * full-field call.  Default 0: exact). */
int svof_set_option(svof_handle* h, const char* name, int value);
/* The CUDA stream (cudaStream_t) every kernel and copy of this handle is ordered on, so a caller that
 * lives on the GPU (halo exchange in multigpu.py) can enqueue its own work in order without a host sync. */
int svof_get_stream(svof_handle* h, void** stream);
/* Block until all device work queued by this handle has finished. */
int svof_synchronize(svof_handle* h);
/* CUDA-event stopwatch on the handle's own stream (the stream every kernel of
 * this handle is launched on): svof_mark records event `slot` (0..7) there,
 * svof_elapsed_ms synchronises on slot b and returns b - a in ms. */
int svof_mark(svof_handle* h, int slot);
int svof_elapsed_ms(svof_handle* h, int slot_a, int slot_b, double* ms);
/* Elapsed device time (ms, CUDA events on the handle's own stream) of the
 * most recent svof_reconstruct + svof_advect pair. */
int svof_last_step_ms(svof_handle* h, double* reconstruct_ms, double* advect_ms);

/* Page-locked host memory for the caller's field storage (what the OpenFOAM
 * adapter pins so svof_step_host's copies run at full PCIe rate). */
int svof_host_alloc(int64_t bytes, void** out);
int svof_host_free(void* p);

/* ---- geometry primitives (unit-test / utility surface) -------------------- */
/* These expose the L1 geometry kernels of the reference on caller-supplied
 * polygons/planes, for known-answer tests and for utilities such as
 * setVofField/exportPlicSurface.H:20-27 that call cutCell directly. */

/* cutFace::calcSubFace (cutFace.C:136-259) on n_polys polygons sharing one
 * vertex count n_verts: pts[n_polys][n_verts][3], plane (normal[3], dist) per
 * polygon.  Out: status[n_polys], centre[3*n_polys], area[3*n_polys]. */
int svof_cut_faces(svof_handle* h, int32_t n_polys, int32_t n_verts, const double* pts,
                   const double* normals, const double* dists, int32_t* status,
                   double* centres, double* areas);

/* cutCell::calcSubCell (cutCell.C:343-542) for a list of mesh cells with given
 * planes.  Out per cell: status, VOF, sub-cell volume, interface centre[3], area[3]. */
int svof_cut_cells(svof_handle* h, int32_t n, const int32_t* cells, const double* normals,
                   const double* dists, int32_t* status, double* vof, double* sub_volume,
                   double* iface_centre, double* iface_area);

/* cutCell::findSignedDistance (cutCell.C:611-799) for a list of mesh cells
 * with given volume fractions and (unit) normals.  Out: status, D, C[3], S[3]. */
int svof_find_signed_distance(svof_handle* h, int32_t n, const int32_t* cells,
                              const double* alphas, const double* normals, int32_t* status,
                              double* dists, double* iface_centre, double* iface_area);

/* cutFace::timeIntegratedFaceFlux (cutFace.C:262-389) for a list of mesh
 * faces.  Out: dVf[n]. */
int svof_face_fluxes(svof_handle* h, int32_t n, const int32_t* faces, const double* normals,
                     const double* dists, const double* Un0, double dt, const double* phi,
                     double* dVf);

/* reconstruction::interface() (reconstruction.C:787-835): the PLIC polygons of the cut cells of the last
 * svof_reconstruct -- per mixed cell with cut status 0, cutCell::interfacePoints (cutCell.C:545-608; evaluated
 * WITHOUT splitWarpedFace, as the reference does).  Face i is points[face_offsets[i] .. face_offsets[i+1]) and belongs
 * to mesh cell cells[i] (ascending).  *n_points / *n_faces always return the sizes; the arrays are filled when
 * points != NULL and the capacities suffice (else SVOF_ERR_CAPACITY).  This is what the plicSurface sampler
 * (src/SimPLIC/sampling) reads.  The angle sort uses atan2, so point coordinates agree between implementations to
 * round-off (1e-13), not bitwise, where two coincident points compete. */
int svof_plic_surface(svof_handle* h, int64_t cap_points, int64_t cap_faces, double* points, int32_t* face_offsets,
                      int32_t* cells, int64_t* n_points, int64_t* n_faces);

/* reconstruction::subCellFaces() (reconstruction.C:838-891): the faces of the SUBMERGED sub-cell of every cut cell of the
 * last svof_reconstruct -- the clipped cell faces, the fully submerged faces and the interface polygon
 * (cutCell::updateSubCellPointsandFaces, cutCell.C:239-290: duplicate points merged at 1e-14, faces oriented away from
 * the sub-cell centre).  Face i is face_points[face_offsets[i] .. face_offsets[i+1]) into points[] and belongs to mesh
 * cell face_cell[i]; cells ascend.  Sizes are always returned; arrays are filled when points != NULL and the capacities
 * suffice.  This is what the reconstructedSubcellFaces sampler reads (sampledReconstructedSubcellFaces.C:97-100). */
int svof_subcell_faces(svof_handle* h, int64_t cap_points, int64_t cap_faces, int64_t cap_face_points, double* points,
                       int32_t* face_offsets, int32_t* face_points, int32_t* face_cell, int64_t* n_points, int64_t* n_faces,
                       int64_t* n_face_points);

/* ---- changing meshes (moving points, refinement) ------------------------------
 * The hooks the reference gets from OpenFOAM's mesh.changing()/moving()/topoChanging(). */

/* Mesh motion without topology change (mesh.moving()): new point positions [3*n_points].  Face/cell geometry, the face
 * flatness (reconstruction::updateFaceFlatness on mesh.changing(), reconstruction.C:643-647) and the tet base points are
 * recomputed; Cf/Sf/C/V may be handed over as in svof_mesh (all four, or all NULL).  The moving-mesh volume scaling of
 * advectionTemplates.C:377-380 acts on a field the update overwrites (alpha1 is recomputed from alpha1.oldTime()), so the
 * new volumes are all the step needs; phi is then the mesh-relative flux, as in the reference's solvers. */
int svof_update_points(svof_handle* h, const double* points, const double* Cf, const double* Sf, const double* C, const double* V);

/* Topology change (dynamicRefineFvMesh refinement / unrefinement; the reference re-does setProcessorPatches on
 * topoChanging, advectionTemplates.C:359-362): every mesh-dependent table of the handle is rebuilt for the new mesh, the
 * parameters are kept.  Fields are NOT mapped here -- OpenFOAM maps its registered fields itself: set alpha (and, for
 * mapAlphaField, the mapped interfaceN/D) afterwards.  Decomposed runs call svof_halo_setup again. */
int svof_update_mesh(svof_handle* h, const svof_mesh* mesh);

/* interfaceN [3*n_cells] and interfaceD [n_cells] as mapped onto the new mesh (the fields `interfaceN`/`interfaceD` the
 * reference registers, reconstruction.C:518-560), for svof_map_alpha_field. */
int svof_set_interface(svof_handle* h, const double* interfaceN, const double* interfaceD);

/* reconstruction::mapAlphaField (reconstruction.C:725-784): in every cell with lower <= alpha <= upper (the
 * dynamicRefineFvMeshCoeffs refinement levels) alpha becomes the volume fraction its (mapped) PLIC plane cuts off,
 * cutCell::calcSubCell without splitWarpedFace; alpha.oldTime() follows.  The caller decides whether to call it
 * (mesh.changing() && mapAlphaField, reconstruction.C:732). */
int svof_map_alpha_field(svof_handle* h, double lower_refine_level, double upper_refine_level);

/* Overset meshes (dynamicOversetFvMesh).  cell_types [n_cells] are cellCellStencil::cellType values of the overset
 * stencil (0 CALCULATED, 1 INTERPOLATED, 2 HOLE, ...): reconstruction::initialize() then lists an interface cell only
 * where the type is CALCULATED (reconstruction.C:649-662).  NULL switches the filter off (any other mesh type).
 * The second overset statement of the path, dVf_ *= faceMask (advectionTemplates.C:383-396, faceMask = the smaller
 * cellMask of the two cells of a face), multiplies by exactly 0 or 1 and is an identity when phi is already masked --
 * which the reference demands of its callers ("Make sure phi_ *= faceMask if overset mesh is used", :370) and its solvers
 * do: pass the masked phi to svof_set_phi.  Call again after the stencil is updated (mesh motion). */
int svof_set_cell_types(svof_handle* h, const int32_t* cell_types);

/* ---- decomposed runs ----------------------------------------------------------
 * Replaces what the reference does across processor patches: the zoneDistribute stencil
 * exchange (reconstruction.C:97-107), syncProcPatches (advection.C:311-393, called once
 * after the geometric fluxes and twice per bounding sweep), setProcessorPatches
 * (advection.C:412-432) and the gMin/gMax loop test (advectionTemplates.C:146-198).
 * Instead of processor patches every rank holds its cells plus `layers` point-neighbour
 * layers of GHOST cells, cut out of the global mesh with an order-preserving renumbering,
 * runs the unchanged single-domain step on that sub-mesh, and ONE exchange per step
 * (ghost alpha <- owning rank, NCCL send/recv on the handle's stream) replaces all of the
 * above.  Owned cells reproduce the single-domain result when the ghost layers cover the
 * dependency radius of a step: 2 + the number of bounding sweeps (default layers =
 * nAlphaBounds + 2).  A mesh with raw processor patches is refused (SVOF_ERR_UNSUPPORTED). */

typedef struct svof_submesh svof_submesh;

/* Cell -> rank map by weighted recursive coordinate bisection (stand-in for decomposePar's
 * scotch, damBreakWithObstacle/system/decomposeParDict:18-20; a cellProcAddressing-derived
 * map from a real scotch run can be used instead).  cell_weight: [n_cells] or NULL. */
int svof_partition_rcb(const svof_mesh* mesh, const double* cell_weight, int32_t n_parts, int32_t* cell_rank_out);

/* Sub-domain of `rank`: owned cells (cell_rank == rank) + `layers` ghost layers.  Local
 * cells/points ascend with their global labels; faces: internal (ascending), the global
 * patches in order (possibly empty), then one extra zeroGradient patch holding the cut
 * faces.  Host only (no CUDA needed). */
int svof_decompose(const svof_mesh* global_mesh, const int32_t* cell_rank, int32_t rank, int32_t layers, svof_submesh** out);
/* View of the sub-mesh (pointers stay owned by the svof_submesh). */
int svof_submesh_mesh(const svof_submesh* s, svof_mesh* out);
/* Addressing: per local cell its global label, owning rank and ghost layer (0 = owned), the
 * local labels of the owned cells, and the global labels of local faces / points
 * (cellProcAddressing / faceProcAddressing / pointProcAddressing).  Any pointer may be NULL. */
int svof_submesh_maps(const svof_submesh* s, int32_t* n_owned, const int32_t** cell_global, const int32_t** cell_owner_rank,
                      const int32_t** cell_layer, const int32_t** owned_local, const int32_t** face_global,
                      const int32_t** point_global);
/* Per local face: the rank owning its global owner cell, and whether it is flipped against the global face. */
int svof_submesh_face_maps(const svof_submesh* s, const int32_t** face_owner_rank, const int32_t** face_flip);
int svof_submesh_free(svof_submesh* s);
const char* svof_decomp_last_error(void);

/* 128-byte NCCL unique id for svof_comm.nccl_unique_id (call on one rank, broadcast). */
int svof_comm_unique_id(void* id128);
/* Collective over the communicator of svof_create: tells the handle which local cells are
 * ghosts (cell_owner_rank[i] != rank) and their global labels (ascending); the ranks
 * exchange the lists once.  From then on svof_advect / svof_step_device end with the ghost
 * refresh, enqueued on the handle's stream (no host synchronisation). */
int svof_halo_setup(svof_handle* h, const int32_t* cell_global, const int32_t* cell_owner_rank);
/* The ghost refresh on its own (after svof_set_alpha with stale ghost values). */
int svof_halo_exchange(svof_handle* h);
/* For callers that only hold the fields of their own cells (an OpenFOAM rank under mpirun: adapter/solveVofEquB200.C):
 * the same plan for faces -- face_owner_rank = rank owning the face's global owner cell, face_flip != 0 where the local
 * face is oriented against the global one (cut faces kept from their neighbour side) -- and one call that refreshes
 * ghost U (cells) and ghost phi (faces) from the owners after svof_set_phi / svof_set_U. */
int svof_halo_setup_faces(svof_handle* h, const int32_t* face_global, const int32_t* face_owner_rank, const int32_t* face_flip);
int svof_halo_exchange_inputs(svof_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* SVOF_H */
