/*
 * drtb.h — C ABI of the B200-native differentiable path tracer hot path.
 *
 * The reference (thalesfm/differentiable-renderer) is a header-only C++17
 * template library with NO FFI layer: its hot path is the triple loop of
 * src/render.cpp:72-86 and everything that loop reaches in include/drt.
 * This header is the boundary a binding for that loop would target: plain C,
 * plain pointers and sizes, integer status codes, no exceptions, no torch
 * types.  Every entry point names the reference interface it replaces.
 *
 * Conventions
 *   - all scene numbers cross the boundary as IEEE double (the reference
 *     instantiates T = double, src/render.cpp:22);
 *   - RGB triples are 3 consecutive doubles;
 *   - images are row-major, row 0 = top (include/drt/camera.hpp:57),
 *     img[(y*W + x)*3 + c];
 *   - gradients are n_params x 3 doubles, UNNORMALISED sums over every sample
 *     of every pixel of seed . d(radiance)/d(param)  (src/render.cpp:78-82 with
 *     the commented `radiance.backward(seed)` enabled);
 *   - every function returns DRTB_OK (0) or a negative DRTB_ERR_* code and
 *     never throws; drtb_last_error() gives the message for the calling ctx;
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with DRTB_ERR_NO_DEVICE.
 */
#ifndef DRTB_H
#define DRTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRTB_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------- */
#define DRTB_OK                 0
#define DRTB_ERR_INVALID       -1   /* bad argument / inconsistent scene       */
#define DRTB_ERR_NO_DEVICE     -2   /* no CUDA device / driver                 */
#define DRTB_ERR_CUDA          -3   /* a CUDA runtime call failed              */
#define DRTB_ERR_UNSUPPORTED   -4   /* valid request this build cannot serve   */
#define DRTB_ERR_NOMEM         -5

/* ---- scene description (flattened include/drt objects) ------------------ */

/* Shape<T> subclasses, include/drt/shape.hpp:37-111 */
#define DRTB_SPHERE 0               /* Sphere: v = {cx, cy, cz, radius}        */
#define DRTB_PLANE  1               /* Plane : v = {nx, ny, nz, offset}; the
                                       normal is used RAW, never normalised
                                       (shape.hpp:58-59)                        */

/* BxDF<T> subclasses, include/drt/bxdf.hpp:56-124 */
#define DRTB_DIFFUSE  0             /* DiffuseBxDF(color), bxdf.hpp:56-83      */
#define DRTB_SPECULAR 1             /* SpecularBxDF(color, exponent),
                                       bxdf.hpp:85-124: half-vector sampling of a
                                       normalised Blinn-Phong lobe, reflect()
                                       (vector.hpp:602-606); direction-dependent
                                       BRDF value, the same two draws per vertex */

typedef struct drtb_prim {
    int32_t type;                   /* DRTB_SPHERE | DRTB_PLANE                */
    int32_t material;               /* index into materials[], -1 = null BxDF
                                       (shape.hpp:14-16, pathtracer.hpp:25-26) */
    int32_t emission;               /* index into params[] of the AreaEmitter's
                                       RGB (emitter.hpp:18), -1 = no emitter   */
    int32_t reserved;
    double  v[4];
} drtb_prim;                        /* 48 bytes; array order == Scene<T> order,
                                       which is the closest-hit tie-break
                                       priority (pathtracer.hpp:78-87)         */

typedef struct drtb_material {
    int32_t type;                   /* DRTB_DIFFUSE | DRTB_SPECULAR            */
    int32_t color;                  /* index into params[] of the albedo RGB
                                       (bxdf.hpp:59-60); shapes sharing one
                                       Vector<T,3,true> share one index         */
    double  exponent;               /* SpecularBxDF::m_exponent (bxdf.hpp:88-91);
                                       ignored for DRTB_DIFFUSE                */
} drtb_material;                    /* 16 bytes                                */

/* Camera<T>, include/drt/camera.hpp:13-37 (after look_at) */
typedef struct drtb_camera {
    int32_t width, height;
    double  vfov;                   /* radians, default 1.3963                 */
    double  eye[3], forward[3], right[3], up[3];
} drtb_camera;

/* Triangle mesh (NEW functionality: the reference has no triangle shape and no
 * acceleration structure).  Semantics are defined to extend the
 * reference's rules: Moller-Trumbore in the precision of the instantiation,
 * both sides hit, acceptance t > 0 with no epsilon, barycentric bounds
 * inclusive (u >= 0, v >= 0, u + v <= 1), det == 0 misses; the normal is the
 * unit geometric normal normalize(cross(v1 - v0, v2 - v0)) used as given, like
 * a plane's.  In scene order the triangles FOLLOW the analytic primitives
 * (triangle i has scene index n_prims + i), so on an exact tie in t the
 * analytic primitive, then the lower triangle index, wins. */
typedef struct drtb_mesh {
    const double*  vertices;   int64_t n_vertices;     /* n_vertices  x 3          */
    const int32_t* indices;    int64_t n_triangles;    /* n_triangles x 3          */
    const int32_t* color;      /* per triangle: params[] index of its DiffuseBxDF
                                  albedo, -1 = null BxDF; NULL = all -1            */
    const int32_t* emission;   /* per triangle: params[] index of its AreaEmitter
                                  RGB, -1 = none; NULL = all -1                    */
} drtb_mesh;

typedef struct drtb_scene {
    const drtb_prim*     prims;      int32_t n_prims;
    const drtb_material* materials;  int32_t n_materials;
    const double*        params;     int32_t n_params;   /* n_params x 3       */
    drtb_camera          camera;
} drtb_scene;

/* ---- render options ------------------------------------------------------ */

#define DRTB_F64   0    /* every path in IEEE double: the parity instantiation  */
#define DRTB_F32   1    /* every path in float (throughput mode; image/gradient
                           accumulators stay double)                            */
#define DRTB_MIXED 2    /* float fast path; any path with a decision closer than
                           its float error bound is re-traced in double          */

#define DRTB_FLAG_IMAGE   1u    /* write the image                              */
#define DRTB_FLAG_GRAD    2u    /* run the adjoint and write the gradients      */
#define DRTB_FLAG_STATS   4u    /* fill drtb_stats (segments, lit paths, ...)   */
#define DRTB_FLAG_NO_BVH  8u    /* test aid: scan every triangle instead of
                                   traversing the BVH (same results, O(N) per ray) */
#define DRTB_FLAG_DETERMINISTIC 16u /* mesh scenes with per-triangle parameters (more
                                   than 8 parameters): sum the gradients in 64-bit
                                   FIXED POINT (resolution 2^-32) instead of floating-
                                   point atomics.  Integer addition is associative, so
                                   the gradients are bit-reproducible run to run and do
                                   not depend on the order the rays finish in; each
                                   contribution is rounded to 2^-32 (|value| < 2^31).
                                   Every other gradient path is deterministic already
                                   and ignores the flag. */

typedef struct drtb_render_opts {
    int32_t  spp;               /* samples per pixel  (args.hpp:32-37 `-n`)     */
    int32_t  min_bounces;       /* Pathtracer::m_min_bounces (`-b`)             */
    double   absorb;            /* Pathtracer::m_absorb      (`-p`)             */
    uint64_t seed;              /* stream selector; 0 reproduces SURVEY KATs    */
    int32_t  precision;         /* DRTB_F64 | DRTB_F32 | DRTB_MIXED             */
    uint32_t flags;             /* DRTB_FLAG_*                                  */
    /* data-parallel shard: image rows are cut into bands of band_rows rows and
       band b belongs to shard (b % shard_count).  shard_count <= 1 = whole
       image.  The shard's rows are written COMPACTLY, in increasing y.        */
    int32_t  shard_index, shard_count, band_rows;
    int32_t  max_depth;         /* vertex-record capacity per path; 0 = default
                                   (min_bounces when absorb == 1, else 64).
                                   Paths still alive at max_depth are cut and
                                   counted in drtb_stats.truncated_paths.       */
    double   seed_scale;        /* adjoint seed = seed_scale * (seed_img ?
                                   seed_img[pixel] : (1,1,1))                    */
    uint64_t adjoint_seed;      /* 0: gradients from the same paths as the image
                                   (the reference's biased mode).  != 0: the
                                   adjoint re-traces with stream `adjoint_seed`
                                   (decorrelated, cf. integrate.hpp:39-52)      */
} drtb_render_opts;

typedef struct drtb_stats {
    uint64_t paths;             /* camera samples traced                        */
    uint64_t segments;          /* ray segments that were intersected           */
    uint64_t lit_paths;         /* paths with non-zero radiance                 */
    uint64_t truncated_paths;   /* paths cut at max_depth                       */
    uint64_t retraced_paths;    /* DRTB_MIXED: paths re-traced in double        */
    uint64_t bvh_nodes;         /* mesh scenes: BVH nodes visited               */
    uint64_t tri_tests;         /* mesh scenes: ray-triangle tests executed     */
    double   kernel_ms;         /* device time of the render kernels (events)   */
} drtb_stats;

typedef struct drtb_ctx drtb_ctx;

/* ---- lifetime ------------------------------------------------------------ */

/* Library/ABI version; callable without a GPU. */
int drtb_abi_version(void);

/* Number of usable CUDA devices (0 without a driver); callable without a GPU. */
int drtb_device_count(void);

/* sizeof() of the ABI structs as this library was compiled, so a binding can
 * verify its mirror: 0 drtb_prim, 1 drtb_material, 2 drtb_camera, 3 drtb_scene,
 * 4 drtb_render_opts, 5 drtb_stats, 6 drtb_mesh; anything else returns 0. */
size_t drtb_struct_size(int which);

/* Create a context on CUDA device `device`.  Replaces nothing in the reference
 * (it has no device); it is the owner of the uploaded scene and scratch. */
int drtb_create(int device, drtb_ctx** out);
void drtb_destroy(drtb_ctx* ctx);

/* Message for the last failing call on ctx (ctx == NULL: last create error). */
const char* drtb_last_error(const drtb_ctx* ctx);

/* ---- scene --------------------------------------------------------------- */

/* Flatten-and-upload: replaces the object graph built at src/render.cpp:26-65
 * (parameters, materials, shapes, Scene<T>, Camera<T>).  Copies everything. */
int drtb_scene_upload(drtb_ctx* ctx, const drtb_scene* scene);

/* Attach a triangle mesh to the uploaded scene (call after drtb_scene_upload;
 * parameter indices refer to that scene's params[]).  Copies the mesh, builds
 * the BVH on the GPU: Morton codes -> radix sort -> binary tree by PLOC
 * (parallel locally-ordered clustering; DRTB_BVH=lbvh selects Karras' radix
 * tree + bottom-up refit instead) -> collapse to an 8-wide compressed BVH
 * (128-byte nodes, quantised child boxes) with <= 3 triangles per leaf.
 * mesh == NULL or n_triangles == 0 detaches the mesh.
 * A later drtb_scene_upload detaches it as well. */
int drtb_mesh_upload(drtb_ctx* ctx, const drtb_mesh* mesh);

/* Device time of the last drtb_mesh_upload's BVH build (all build kernels +
 * the radix sort), in milliseconds; 0 if no mesh is attached. */
double drtb_mesh_build_ms(const drtb_ctx* ctx);

/* Overwrite parameter values only (n_params x 3 doubles); the cheap call an
 * optimisation loop makes between renders. */
int drtb_set_params(drtb_ctx* ctx, const double* params, int32_t n_params);

/* Same from DEVICE memory, enqueued on `stream` (ordered with the renders the caller enqueues there;
 * asynchronous): an optimisation loop whose update step runs on the GPU never returns to the host. */
int drtb_set_params_device(drtb_ctx* ctx, const double* d_params, int32_t n_params, void* stream);

/* Rows of the image that shard (index, count, band_rows) owns, for sizing the
 * compact shard buffers; callable without a GPU. */
int32_t drtb_shard_rows(int32_t height, int32_t shard_index, int32_t shard_count,
                        int32_t band_rows);

/* ---- the hot path -------------------------------------------------------- */

/* Forward render + adjoint with HOST buffers: replaces the pixel loop of
 * src/render.cpp:72-86 (cam.sample -> tracer.trace -> accumulate ->
 * radiance.backward(seed)), i.e. Camera::sample (camera.hpp:51-60),
 * Pathtracer::trace/raycast/scatter (pathtracer.hpp:72-136), DiffuseBxDF
 * (bxdf.hpp:63-79), AreaEmitter (emitter.hpp:20-21) and the reverse tape
 * (vector.hpp:120-318, 418-557).
 *   seed_img : NULL or shard_rows*W*3 doubles, per-pixel adjoint seed
 *   img      : shard_rows*W*3 doubles (may be NULL without DRTB_FLAG_IMAGE)
 *   grad     : n_params*3 doubles     (may be NULL without DRTB_FLAG_GRAD);
 *              OVERWRITTEN with this call's sums (the caller adds them to
 *              VariableNode::m_grad, vector.hpp:185-188)
 *   stats    : NULL or filled
 * Blocking.  Host->device and device->host copies happen inside the call.
 * A PINNED `img` (cudaHostAlloc / cudaHostRegister) is written by the render
 * kernel itself, pixel by pixel under the compute, and no image copy follows
 * the kernel (analytic scenes, DRTB_F64 / DRTB_F32); pageable memory works as
 * well and costs one device->host copy after the kernel.  Same bits. */
int drtb_render(drtb_ctx* ctx, const drtb_render_opts* opts,
                const double* seed_img, double* img, double* grad,
                drtb_stats* stats);

/* Pinned host memory for drtb_render's buffers without the CUDA headers
 * (cudaHostAlloc, portable: every GPU of the box can address it).  An image
 * allocated here takes the kernel-written path described above.
 * drtb_host_free(NULL) is a no-op. */
int drtb_host_alloc(size_t bytes, void** ptr);
int drtb_host_free(void* ptr);

/* Same, DEVICE buffers on ctx's device, enqueued on `stream` (a cudaStream_t
 * passed as void*; NULL = the legacy default stream).  Asynchronous: returns
 * after the launches; the caller synchronises the stream.  d_stats is NULL or
 * a device drtb_stats, of which the kernels fill segments, lit_paths,
 * truncated_paths, bvh_nodes and tri_tests (paths, retraced_paths and
 * kernel_ms are host-side figures that only drtb_render fills).
 * ONE STREAM AT A TIME PER CONTEXT: every render of a ctx shares the ctx's
 * scratch (task counter, gradient partials, lit-path rings, wavefront
 * buffers), so two renders of the same ctx must not be in flight on different
 * streams at once -- enqueue them on one stream, or order the streams with an
 * event.  drtb_render / drtb_set_params run on the ctx's own private stream
 * and block until done, so they are always ordered with each other; mixing
 * them with *_device calls still in flight on another stream is the caller's
 * to order (synchronise that stream first). */
int drtb_render_device(drtb_ctx* ctx, const drtb_render_opts* opts,
                       const double* d_seed_img, double* d_img, double* d_grad,
                       drtb_stats* d_stats, void* stream);

/* Size every scratch buffer a render with these options needs, set the kernel
 * attributes and load the kernels it will launch, so that the first
 * drtb_render_device with them is as fast as the second (drtb_render does the
 * same by itself before it starts the timer behind drtb_stats.kernel_ms).
 * Optional; blocking. */
int drtb_reserve(drtb_ctx* ctx, const drtb_render_opts* opts);

/* ---- multi-GPU: the image all-gather fused into the render ----------------
 * The reference renders in one process (src/render.cpp:72-86 fills one img[]).
 * With the image rows sharded over the GPUs of one NVSwitch box (shard_index /
 * shard_count / band_rows), every GPU normally ends up with its own rows only
 * and a gather has to follow.  Instead the render kernel can store each pixel,
 * as it is finished, into the FULL image (H*W*3 doubles, row = image row, row
 * 0 = top) of every GPU of the job: peer stores over NVLink, hidden behind the
 * tracing.  full_images[p] are device pointers valid in this process -- this
 * GPU's own buffer and the peers' buffers opened with drtb_ipc_open (or any
 * peer-accessible allocation).  n = 0 switches it off.  While it is on,
 * DRTB_FLAG_IMAGE renders (analytic scenes) fill the peers' images and d_img
 * may be NULL.  The stores are complete when the kernel is; a rank may read
 * its full image once every rank's render has finished (the gradient
 * all-reduce that follows the render on each rank's stream orders that).
 * The opposite hazard is the caller's as well: a rank's NEXT peer-filling
 * render stores straight into the other ranks' full images, so a cross-rank
 * barrier or collective must separate every rank's last read of its full
 * image from the next such render on any rank (or the caller double-buffers
 * the full images and alternates drtb_set_image_peers). */
int drtb_set_image_peers(drtb_ctx* ctx, double* const* full_images, int32_t n);

/* ---- multi-GPU: the gradient sum without a collective library ----------------
 * north_star asks for "one NCCL allreduce" of each GPU's parameter gradients.  For the Cornell box that is 96
 * bytes, and its cost is pure latency (~0.1 ms per render at 8 GPUs, a third of a 256 x 256 render).  Since every
 * GPU of the box can store into every other GPU's memory, the library can do the sum itself: every rank allocates
 * an exchange buffer of drtb_grad_exchange_bytes(n, n_params) bytes, ZERO-INITIALISED (cudaMemset), shares it
 * (drtb_ipc_alloc / drtb_ipc_open across processes, plain peer access within one), and hands the rank-ordered list
 * to its context.  From then on every render with DRTB_FLAG_GRAD ends with a one-block kernel that stores this
 * rank's gradients into all buffers, waits for the other ranks' and adds them in rank order: `grad` holds the SUM
 * over the job, bit-identical on every rank, and that kernel is also what orders "every rank's image stores have
 * landed" (drtb_set_image_peers).  All ranks must issue the same sequence of such renders (it is a collective);
 * a rank that never arrives poisons the sums with NaN after ~2 s instead of hanging the GPUs.  Up to 4096
 * gradient scalars (mesh scenes with per-triangle parameters keep the collective library).  n = 0 switches it off. */
size_t drtb_grad_exchange_bytes(int32_t n_ranks, int32_t n_params);
int drtb_set_grad_peers(drtb_ctx* ctx, void* const* exchange, int32_t n, int32_t rank);

/* Device memory that another process on the same box can map (CUDA IPC, one
 * process per GPU).  drtb_ipc_alloc: cudaMalloc on ctx's device, zero-filled, + its
 * DRTB_IPC_HANDLE_BYTES-byte handle, to be sent to the peers by the host's own
 * means; drtb_ipc_open: map a peer's handle into this process (peer access is
 * enabled on demand); drtb_ipc_close / drtb_ipc_free undo them. */
#define DRTB_IPC_HANDLE_BYTES 64
int drtb_ipc_alloc(drtb_ctx* ctx, size_t bytes, void** d_ptr, void* handle);
int drtb_ipc_open(drtb_ctx* ctx, const void* handle, void** d_ptr);
int drtb_ipc_close(drtb_ctx* ctx, void* d_ptr);
int drtb_ipc_free(drtb_ctx* ctx, void* d_ptr);

/* ---- multi-GPU in one process -------------------------------------------------
 * The GPUs of one NVSwitch box behind ONE handle: what a C or C++ caller of the drop-in headers uses to reach
 * them (drt::RenderOptions::devices, `build/render --gpus N`); a job with one process per GPU uses the shard_*
 * fields of drtb_render_opts with drtb_set_image_peers / drtb_set_grad_peers instead.  Every call mirrors its
 * single-device namesake; the scene is uploaded to (and a mesh's BVH built on) every device.
 * drtb_multi_render replaces the pixel loop of src/render.cpp:72-86 like drtb_render does:
 *   opts     : shard_index / shard_count are set by the library (band b of band_rows rows -> device b mod n;
 *              band_rows <= 0 means 8)
 *   seed_img : NULL or the FULL H*W*3 per-pixel adjoint seed
 *   img      : the FULL H*W*3 image.  Analytic scenes: every device's kernel stores its pixels into the full
 *              image on the first device over NVLink, which then leaves in one copy; mesh scenes (and devices
 *              without peer access): each device's bands are copied to their rows
 *   grad     : n_params*3 doubles, the SUM over the devices (added on the host in device order)
 *   stats    : NULL or the totals over the devices; kernel_ms = the slowest device's
 * Blocking; one host thread at a time per handle. */
typedef struct drtb_multi drtb_multi;
int drtb_multi_create(const int* devices, int32_t n, drtb_multi** out);      /* 1 <= n <= 8, distinct devices */
void drtb_multi_destroy(drtb_multi* m);
const char* drtb_multi_last_error(const drtb_multi* m);                       /* m == NULL: last create error */
int32_t drtb_multi_device_count(const drtb_multi* m);
int drtb_multi_scene_upload(drtb_multi* m, const drtb_scene* scene);
int drtb_multi_mesh_upload(drtb_multi* m, const drtb_mesh* mesh);
int drtb_multi_set_params(drtb_multi* m, const double* params, int32_t n_params);
int drtb_multi_render(drtb_multi* m, const drtb_render_opts* opts, const double* seed_img, double* img, double* grad,
                      drtb_stats* stats);

/* drtb_render plus the PER-PIXEL gradient image of ONE parameter (the figure of
 * README.md:138-145, "gradients of the pixel colors with respect to the
 * parameter controlling the color of the left wall"): the same adjoint sweep,
 * with parameter `param`'s contributions additionally summed per pixel.
 *   grad_img : shard_rows*W*3 doubles,
 *              grad_img[(y*W + x)*3 + c] = sum over the pixel's samples of
 *              seed_c * d radiance_c / d params[param][c]   (seed as in
 *              drtb_render; with seed_scale = 1/spp and no seed_img this is
 *              d pixel_c / d param_c).  Summed over all pixels it equals
 *              grad[param].
 * Needs DRTB_FLAG_GRAD; img / grad as in drtb_render. */
int drtb_render_grad_image(drtb_ctx* ctx, const drtb_render_opts* opts, int32_t param,
                           const double* seed_img, double* img, double* grad,
                           double* grad_img, drtb_stats* stats);

/* Same with DEVICE buffers on `stream` (asynchronous, see drtb_render_device). */
int drtb_render_grad_image_device(drtb_ctx* ctx, const drtb_render_opts* opts, int32_t param,
                                  const double* d_seed_img, double* d_img, double* d_grad,
                                  double* d_grad_img, drtb_stats* d_stats, void* stream);

/* Batch of explicit rays: replaces Pathtracer<T>::trace(scene, orig, dir)
 * (pathtracer.hpp:121-136) called from user code, plus radiance.backward().
 *   orig, dir : n x 3 doubles (dir is used as given, as the reference does)
 *   keys      : n stream keys; ray i draws u(keys[i], slot) with the first
 *               scatter draw at slot 2 (slots 0,1 are the camera's)
 *   radiance  : n x 3 doubles
 *   jac       : NULL or n x n_params x 3 doubles, d radiance_c / d param_{k,c}
 *               (channels never mix) so that
 *               backward(g) == param_k.grad[c] += g[c] * jac[i][k][c]        */
int drtb_trace_rays(drtb_ctx* ctx, const drtb_render_opts* opts, int64_t n,
                    const double* orig, const double* dir, const uint64_t* keys,
                    double* radiance, double* jac);

/* ---- measurement helpers -------------------------------------------------- */

/* Dependent-free FMA issue-rate micro-benchmark on ctx's device: the roofline
 * denominator for this compute-bound path (MEASURED_PEAKS.json has only HBM
 * and tensor numbers).  precision = DRTB_F64 | DRTB_F32; result in TFLOP/s
 * (FMA = 2 FLOP). */
int drtb_fma_peak(drtb_ctx* ctx, int32_t precision, double* tflops);

/* How the render kernels cut n_units units of work (warp tasks, or pixels when regen != 0: the path-regenerating
 * kernel of Russian-roulette renders) into the chunks that resident warps claim from a global counter:
 * out = { units per big chunk, units per small chunk, number of big chunks, number of chunks }.  Chunk c covers
 * units [c * big, (c + 1) * big) for c < n_big and [n_big * big + (c - n_big) * small, ... + small) clipped to
 * n_units after that.  Host arithmetic, callable without a GPU (the tests check the cover). */
int drtb_chunk_plan(int64_t n_units, int32_t spp, int64_t resident_warps, int32_t regen, int64_t out[4]);

/* Number of kernels this context has launched so far (bench `gpu_launches`). */
uint64_t drtb_launch_count(const drtb_ctx* ctx);

/* The counter-based sample stream, exposed so hosts and tests can reproduce
 * it: k = splitmix64(key * 0x100000001B3 + slot) % 2147483647, the value the
 * reference's random::uniform() (random.hpp:7-10) would get from rand().
 * key = seed * 0x9E3779B97F4A7C15 + (y*W + x)*spp + i.  Callable without GPU. */
uint32_t drtb_stream_draw(uint64_t key, uint32_t slot);

#ifdef __cplusplus
}
#endif
#endif /* DRTB_H */
