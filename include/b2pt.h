/*
 * b2pt.h -- C ABI of the B200-native wavefront path tracer.
 *
 * This is the drop-in boundary for ONE hot path of nkkk98/MyGPURaytracer: the
 * per-iteration wavefront loop behind
 *
 *     void pathtraceInit(Scene*);                          apps/src/pathtrace.h:7
 *     void pathtraceFree();                                apps/src/pathtrace.h:8
 *     void pathtrace(uchar4* pbo, int frame, int iter);    apps/src/pathtrace.h:9
 *     PerformanceTimer& timer();                           apps/src/pathtrace.h:6
 *     void sendToGPU(uchar4* pbo, int iter);               apps/src/pathtrace.h:10
 *
 * Everything here is plain C: PODs, pointers and sizes.  No C++ types, no
 * torch types, no glm.  Functions return 0 (B2PT_OK) or a negative error code
 * and never call exit() (the reference's checkCUDAError does,
 * apps/src/pathtrace.cu:46-64).  b2pt_last_error() returns the text of the
 * last failure on the calling thread.
 *
 * Layout conventions
 *   - matrices are 16 floats, column-major (m[col*4+row]), the layout of
 *     glm::mat4 in the reference's Geom (apps/src/sceneStructs.h:50-70);
 *   - vectors are tightly packed floats;
 *   - images are W*H*3 floats, index = x + y*W (apps/src/pathtrace.cu:255).
 */
#ifndef B2PT_H_
#define B2PT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2PT_ABI_VERSION 2 /* 2: B2ptScene::face_material / material_textures, B2ptLoadOverrides::per_face_materials */

/* ---- error codes ------------------------------------------------------ */
enum {
  B2PT_OK = 0,
  B2PT_ERR_INVALID = -1,  /* bad argument / inconsistent scene           */
  B2PT_ERR_CUDA = -2,     /* a CUDA runtime call failed                  */
  B2PT_ERR_NOMEM = -3,    /* host or device allocation failed            */
  B2PT_ERR_IO = -4,       /* scene / OBJ / texture file problem          */
  B2PT_ERR_STATE = -5,    /* call made in the wrong state                */
  B2PT_ERR_RANGE = -6     /* index / capacity out of range               */
};

/* ---- scene PODs -------------------------------------------------------- */

/* enum GeomType, apps/src/sceneStructs.h:10-15 (same numeric values). */
enum { B2PT_SPHERE = 0, B2PT_CUBE = 1, B2PT_TRIANGLE = 2, B2PT_OBJ = 3 };

/* struct Material, apps/src/sceneStructs.h:72-82 (same 44-byte layout). */
typedef struct B2ptMaterial {
  float color[3];
  float specular_exponent;
  float specular_color[3];
  float has_reflective;
  float has_refractive;
  float index_of_refraction;
  float emittance;
} B2ptMaterial;

/* struct Texture, apps/src/sceneStructs.h:36-48.  Raw 8-bit interleaved
 * texels, row 0 first (the reference flips images at load,
 * apps/src/scene.cpp:133; the bytes here are post-flip). */
typedef struct B2ptTexture {
  int32_t width;
  int32_t height;
  int32_t channels; /* 0 = "no texture" (Texture() ctor) */
  int32_t reserved;
  const uint8_t* texels; /* width*height*channels bytes, host memory */
} B2ptTexture;

/* struct Geom, apps/src/sceneStructs.h:50-70, minus the fields the hot path
 * never reads (translation/rotation/scale/minPos/maxPos).  Faces and
 * textures are referenced by index instead of by device pointer. */
typedef struct B2ptGeom {
  int32_t type;        /* B2PT_SPHERE / B2PT_CUBE / B2PT_TRIANGLE / B2PT_OBJ */
  int32_t material_id; /* Geom::materialid */
  float transform[16];
  float inverse_transform[16];
  float inv_transpose[16];
  int32_t face_begin; /* first face of this geom in the scene face arrays */
  int32_t face_count; /* Geom::faceSize */
  int32_t tex_kd;     /* index into B2ptScene::textures or -1 */
  int32_t tex_ks;
  int32_t tex_bump;
  int32_t tex_ke;
} B2ptGeom;

/* struct Camera, apps/src/sceneStructs.h:84-93 (same 84-byte layout).
 * These are the values pathtrace() actually reads, i.e. AFTER the orbit
 * recompute of apps/src/main.cpp:222-240; `right` is NOT normalised there. */
typedef struct B2ptCamera {
  int32_t resolution[2];
  float position[3];
  float look_at[3];
  float view[3];
  float up[3];
  float right[3];
  float fov[2];
  float pixel_length[2];
} B2ptCamera;

/* A whole scene.  Faces of all OBJ geoms live in two flat arrays:
 *   face_pos: 9 floats per face  (Face::v0/v1/v2 .position)
 *   face_uv : 6 floats per face  (Face::v0/v1/v2 .texcoord)
 * which is the only part of the reference's 156-byte Face the path reads
 * (apps/src/intersections.h:216-249). */
typedef struct B2ptScene {
  int32_t n_geoms;
  int32_t n_materials;
  int32_t n_textures;
  int32_t n_faces;
  const B2ptGeom* geoms;
  const B2ptMaterial* materials;
  const B2ptTexture* textures;
  const float* face_pos;
  const float* face_uv;
  B2ptCamera camera;
  int32_t trace_depth; /* RenderState::traceDepth */
  int32_t iterations;  /* RenderState::iterations (informational) */
  /* Per-face materials, optional (both NULL = the reference's behaviour).  The
   * reference reads tinyobj's mesh.material_ids[f] and discards it
   * (apps/src/scene.cpp:121-122): every face of an OBJ is shaded with the one
   * material the loader appends per OBJ from MTL material 0 (:68,134,221-231).
   *   face_material[f]         scene material of face f (n_faces entries, every
   *                            entry in [0, n_materials))
   *   material_textures[4*m+k] texture index (or -1) of scene material m for
   *                            k = 0 kd, 1 ks, 2 bump, 3 ke (4*n_materials
   *                            entries): with it an OBJ face takes its four maps
   *                            from its material instead of from its geom */
  const int32_t* face_material;
  const int32_t* material_textures;
} B2ptScene;

/* ---- options ------------------------------------------------------------ */

/* Trigonometry used by the cosine-hemisphere / lens sampling
 * (apps/src/interactions.h:42-43, apps/src/pathtrace.cu:238):
 *   NATIVE   - CUDA libdevice sinf/cosf/pow: what the reference's kernels call;
 *   PORTABLE - a fixed +,* only polynomial shared bit-for-bit with the CPU
 *              oracle (<= 2 ulp from NATIVE); used by the parity tests. */
enum { B2PT_TRIG_NATIVE = 0, B2PT_TRIG_PORTABLE = 1 };

/* RNG keying.  SLOT reproduces the reference: the shader seeds with the
 * array slot after sort+compaction (apps/src/pathtrace.cu:467).  PIXEL keys
 * a counter on (pixel, iteration, depth) and cannot match the reference
 * per iteration (only the converged image). */
enum { B2PT_RNG_SLOT = 0, B2PT_RNG_PIXEL = 1 };

typedef struct B2ptOptions {
  uint32_t struct_size;       /* sizeof(B2ptOptions), for ABI growth       */
  int32_t device;             /* CUDA device ordinal                        */
  int32_t antialiasing;       /* ANTIALIASING, pathtrace.cu:39   (def 1)    */
  int32_t depth_of_field;     /* DEPTH_OF_FIELD, pathtrace.cu:36 (def 0)    */
  float lens_radius;          /* pathtrace.cu:279 (def 0.8)                 */
  float focal_distance;       /* pathtrace.cu:280 (def 11)                  */
  int32_t sort_by_material;   /* SORT_BY_MATERIAL, pathtrace.cu:38 (def 1)  */
  int32_t cache_first_bounce; /* intended semantics of pathtrace.cu:586-610;
                                 only honoured when AA and DOF are off (def 0) */
  int32_t trig_mode;          /* B2PT_TRIG_*  (def NATIVE)                  */
  int32_t rng_mode;           /* B2PT_RNG_*   (def SLOT)                    */
  int32_t use_bvh;            /* 1 = LBVH traversal, 0 = brute force (def 1)*/
  int32_t record_stages;      /* 1 = keep per-depth stage dumps of the next
                                 b2pt_render call readable (def 0)          */
  int32_t use_graph;          /* 1 = replay the iteration as a CUDA graph   */
  int32_t concurrent_contexts; /* how many contexts render on this GPU at the
                                 same time (spp sharding inside one GPU, each
                                 on its own stream).  > 1 sizes the persistent
                                 grids to a share of the SMs so that the
                                 contexts' kernels co-reside: lower latency
                                 per context is traded for throughput (def 1) */
  int32_t persistent_host_albedo; /* b2pt_pathtrace / b2pt_pipe_pathtrace copy the
                                 albedo AOV to the host on EVERY call, as the
                                 reference does (apps/src/pathtrace.cu:666-668)
                                 (def 0).  1 = the caller asserts that the
                                 albedo_host buffer it passes is persistent and
                                 that nobody else writes it between calls (the
                                 reference's scene->state.albedo is such a
                                 buffer): the copy is then made only when the
                                 AOV changed (iteration 1, reset, camera) or
                                 the pointer is new                            */
  int32_t reserved[6];
} B2ptOptions;

/* Fills `opt` with the defaults above (the reference's compile-time macros). */
void b2pt_default_options(B2ptOptions* opt);

/* ---- context ------------------------------------------------------------- */
typedef struct B2ptCtx B2ptCtx;

/* pathtraceInit (apps/src/pathtrace.cu:130-194): uploads the scene as SoA
 * buffers, builds the LBVH of every OBJ geom on the GPU, allocates the path
 * state, hit records and the zeroed accumulation / albedo images. */
int b2pt_create(const B2ptScene* scene, const B2ptOptions* opt, B2ptCtx** out);

/* Another context on the parent's GPU that SHARES the parent's scene on the
 * device (geoms, materials, BVHs, triangles, textures: read only for every
 * kernel) and owns only its path state and accumulators.  `scene` must be the
 * scene the parent was created from (camera, resolution and depth are taken
 * from it).  Contexts that render the same scene at once (b2pt_pipe_*, spp
 * sharding inside a GPU) then walk ONE BVH, which stays resident in L2, instead
 * of one copy each.  The shared data lives until the last context using it is
 * destroyed; parent and children may be destroyed in any order. */
int b2pt_create_shared(B2ptCtx* parent, const B2ptScene* scene, const B2ptOptions* opt, B2ptCtx** out);

/* pathtraceFree (apps/src/pathtrace.cu:196-223).  NULL is accepted. */
void b2pt_destroy(B2ptCtx* ctx);

/* Replace the camera (the reference re-runs pathtraceInit when the camera
 * moves, apps/src/main.cpp:222-248) and zero the accumulators. */
int b2pt_set_camera(B2ptCtx* ctx, const B2ptCamera* cam);

/* Zero the accumulation and albedo images (what Free+Init does to them). */
int b2pt_reset_accum(B2ptCtx* ctx);

/* Render iterations iter_first, iter_first+iter_stride, ... (iter_count of
 * them) and add each into the device accumulation image:
 *     image[pixel] += color * PI          (apps/src/pathtrace.cu:501-510)
 * One call with (iter, 1, 1) is the device part of one reference pathtrace()
 * call.  iter_stride > 1 is the samples-per-pixel sharding used across GPUs:
 * rank r of R renders (r+1, count, R).  Asynchronous on the context's stream. */
int b2pt_render(B2ptCtx* ctx, int32_t iter_first, int32_t iter_count, int32_t iter_stride);

/* Wait for everything queued on the context's stream. */
int b2pt_sync(B2ptCtx* ctx);

/* The D2H half of pathtrace() (apps/src/pathtrace.cu:663-668): copy the
 * un-normalised running sum and the iteration-1 albedo to host memory
 * (W*H*3 floats each; either may be NULL).  Synchronous. */
int b2pt_read_accum(B2ptCtx* ctx, float* image_host, float* albedo_host);

/* Exactly one reference pathtrace(pbo, frame, iter) call: render `iter`, then
 * copy image and albedo to the host buffers (scene->state.image / .albedo).
 * Both are copied on every call.  With B2ptOptions.persistent_host_albedo = 1
 * the copy into `albedo_host` is skipped when the previous call already wrote
 * the current albedo to the SAME pointer (the AOV only changes on iteration 1
 * and on a reset); b2pt_read_accum always copies. */
int b2pt_pathtrace(B2ptCtx* ctx, int32_t iter, float* image_host, float* albedo_host);

/* ---- pipelined pathtrace() --------------------------------------------------------
 * The reference's host asks for ONE iteration per call, with consecutive
 * iteration numbers (pathtrace(pbo, 0, ++iteration), apps/src/main.cpp:255), and
 * reads scene->state.image after every call (apps/src/pathtrace.cu:663-668).
 * A B2ptPipe serves exactly that contract with `lanes` contexts on one GPU:
 * while the host consumes iteration i (merge into the running sum + copy to the
 * host), the lanes already render i+stride, i+2*stride, ...  Each iteration is
 * rendered once and merged in call order; the host image is bit-identical to
 * b2pt_pathtrace() on a single context.  A call that does not continue the
 * sequence (restart, another stride) drops the speculated iterations, adopts
 * the new stride (iter - previous iter) and starts over; nothing is lost but
 * the overlap.  (csrc/pipe.cu) */
typedef struct B2ptPipe B2ptPipe;

/* pathtraceInit for a pipelined host.  `opt` as for b2pt_create
 * (concurrent_contexts is set to `lanes`; record_stages is refused).
 * 1 <= lanes <= 16; bench.py measures 6 (4 / 5 / 6 lanes: 2559 / 2604 / 2610 Mpaths/s on one B200). */
int b2pt_pipe_create(const B2ptScene* scene, const B2ptOptions* opt, int32_t lanes, B2ptPipe** out);

/* pathtraceFree.  NULL is accepted. */
void b2pt_pipe_destroy(B2ptPipe* pipe);

/* One reference pathtrace(pbo, frame, iter) call: on return image_host holds
 * the running sum including iteration `iter`, albedo_host the iteration-1
 * albedo AOV (every call; with persistent_host_albedo = 1 only when it changed
 * or the pointer is new; either may be NULL).  Synchronous for the caller; the
 * lanes keep rendering ahead. */
int b2pt_pipe_pathtrace(B2ptPipe* pipe, int32_t iter, float* image_host, float* albedo_host);

/* Zero the running sum, the albedo and every lane (Free + Init of
 * apps/src/main.cpp:245-248); with `cam` != NULL also replace the camera. */
int b2pt_pipe_reset(B2ptPipe* pipe, const B2ptCamera* cam);

/* The running sum / the albedo AOV on the device (W*H*3 floats), valid after
 * a b2pt_pipe_pathtrace call: the inputs of b2pt_tonemap_rgba8 and of a
 * device-side denoiser. */
float* b2pt_pipe_device_image(B2ptPipe* pipe);
float* b2pt_pipe_device_albedo(B2ptPipe* pipe);

/* sendImageToPBO / sendDenosiedImageToPBO for a pipelined host: as
 * b2pt_tonemap_rgba8, `src_dev` NULL = the running sum.  Runs on the pipe's
 * copy stream (never behind the iterations the lanes render ahead) and
 * returns when rgba8_dev is written, so the caller may unmap / draw the PBO
 * and overwrite `src_dev` at once (apps/src/main.cpp:270-271). */
int b2pt_pipe_tonemap_rgba8(B2ptPipe* pipe, const float* src_dev, int32_t iter, uint8_t* rgba8_dev);

/* Introspection: number of lanes, the context of lane k (statistics, BVH
 * info, tonemap), kernel launches of all lanes + merges, and how often a call
 * did not continue the predicted sequence. */
int32_t b2pt_pipe_lanes(B2ptPipe* pipe);
B2ptCtx* b2pt_pipe_lane(B2ptPipe* pipe, int32_t k);
int64_t b2pt_pipe_launch_count(B2ptPipe* pipe);
int64_t b2pt_pipe_misses(B2ptPipe* pipe);

/* timer().getGpuElapsedTimeForPreviousOperation() for a pipelined host
 * (apps/src/main.cpp:263): milliseconds of the depth loop
 * (apps/src/pathtrace.cu:583-653) of the iteration the last
 * b2pt_pipe_pathtrace call consumed. */
float b2pt_pipe_last_loop_ms(B2ptPipe* pipe);

/* ---- pathtrace() over several GPUs ----------------------------------------------
 * Iterations are independent (the only state they share is the additive
 * accumulator, apps/src/pathtrace.cu:508), so G GPUs render G consecutive
 * iterations at once.  One FRAME = iterations first .. first+G-1, member g renders
 * first+g into its own zeroed image; `lanes` contexts per GPU render the next
 * frames ahead (as a B2ptPipe does on one GPU).  A frame is combined by ONE kernel
 * per member over NVLink peer memory (csrc/multi.cu, k_frame_reduce): member g
 * owns pixel slice g of the running sum, reads slice g of every member's image,
 * adds them IN ITERATION ORDER, clears what it consumed and forwards the new sum
 * into member 0's whole image.  The additions are the reference's sequential
 * image[pixel] += color*PI, so the result is bit-identical to one GPU rendering
 * the same iterations one after the other.
 */
#define B2PT_MAX_MEMBERS 16

/* (a) One process, several GPUs (the reference's host is one process,
 * apps/src/main.cpp).  `devices` = n_devices CUDA ordinals (NULL: 0..n-1; an
 * ordinal may repeat: those members share the device).  The devices must have
 * peer access to each other (NVLink / NVSwitch).  This is pathtraceInit. */
typedef struct B2ptMulti B2ptMulti;
int b2pt_multi_create(const B2ptScene* scene, const B2ptOptions* opt, int32_t n_devices, const int32_t* devices,
                      int32_t lanes, B2ptMulti** out);
void b2pt_multi_destroy(B2ptMulti* multi); /* pathtraceFree; NULL is accepted */
/* G reference pathtrace() calls at once: renders iterations first_iter ..
 * first_iter+G-1 (consecutive frames are predicted and rendered ahead) and
 * returns the running sum including all of them in image_host, the iteration-1
 * albedo AOV in albedo_host (either may be NULL).  Every member copies its own
 * slice of the image over its own PCIe link. */
int b2pt_multi_pathtrace(B2ptMulti* multi, int32_t first_iter, float* image_host, float* albedo_host);
int b2pt_multi_reset(B2ptMulti* multi, const B2ptCamera* cam);
int32_t b2pt_multi_members(B2ptMulti* multi);
float* b2pt_multi_device_image(B2ptMulti* multi); /* the running sum on member 0's device */
int64_t b2pt_multi_launch_count(B2ptMulti* multi);

/* (b) One process per GPU (torch.distributed, MPI): a B2ptShard per rank.  Device
 * pointers are exchanged once as CUDA IPC handles: every rank exports a blob of
 * b2pt_shard_export_size() bytes, the host all-gathers the blobs (rank order) and
 * hands them to b2pt_shard_connect.  A frame is three calls made by EVERY rank in
 * the same order, with a cross-rank barrier queued by the host on
 * b2pt_shard_stream() between them (NCCL: a 4-byte all-reduce on that stream):
 *
 *     b2pt_shard_frame_begin(s, first_iter);   my contribution of the frame is complete
 *     <barrier on b2pt_shard_stream(s)>        ... and so is everybody else's
 *     b2pt_shard_frame_reduce(s);              k_frame_reduce on my slice
 *     <barrier on b2pt_shard_stream(s)>        every slice consumed, member 0's image complete
 *     b2pt_shard_frame_end(s, image_host, albedo_host);
 *
 * frame_end copies the running sum / the albedo AOV to host memory on rank 0 (the
 * pointers are ignored elsewhere; NULL = keep it on the device) and re-arms the
 * lane.  A rank that is given host pointers waits for the copy; the others return
 * at once and are throttled by frame_begin (at most `lanes` frames ahead).
 *
 * The plain alternative to begin/barrier/reduce/barrier, for hosts that prefer a
 * library collective: after frame_begin, reduce b2pt_shard_frame_image(s) (W*H*3
 * floats, sum, to rank 0, in place) on b2pt_shard_stream(s) -- e.g. ncclReduce --
 * then call b2pt_shard_frame_merge(s) and frame_end.  The image then agrees with
 * the single-GPU one to float summation order instead of bit for bit. */
typedef struct B2ptShard B2ptShard;
int b2pt_shard_create(const B2ptScene* scene, const B2ptOptions* opt, int32_t rank, int32_t world, int32_t lanes,
                      B2ptShard** out);
void b2pt_shard_destroy(B2ptShard* shard);
int64_t b2pt_shard_export_size(B2ptShard* shard);
int b2pt_shard_export(B2ptShard* shard, void* blob, int64_t bytes);
int b2pt_shard_connect(B2ptShard* shard, const void* blobs_in_rank_order, int64_t bytes_per_member);
int b2pt_shard_frame_begin(B2ptShard* shard, int32_t first_iter);
void* b2pt_shard_stream(B2ptShard* shard); /* cudaStream_t of the frame protocol */
int b2pt_shard_frame_reduce(B2ptShard* shard);
float* b2pt_shard_frame_image(B2ptShard* shard);
int b2pt_shard_frame_merge(B2ptShard* shard);
int b2pt_shard_frame_end(B2ptShard* shard, float* image_host, float* albedo_host);
int b2pt_shard_reset(B2ptShard* shard, const B2ptCamera* cam);
int b2pt_shard_sync(B2ptShard* shard);
float* b2pt_shard_device_image(B2ptShard* shard); /* rank 0: the running sum; NULL elsewhere */
float* b2pt_shard_device_slice(B2ptShard* shard, int64_t* offset_floats, int64_t* len_floats);
int32_t b2pt_shard_lanes(B2ptShard* shard);
B2ptCtx* b2pt_shard_lane(B2ptShard* shard, int32_t k);
B2ptShard* b2pt_multi_member(B2ptMulti* multi, int32_t g);
int64_t b2pt_shard_launch_count(B2ptShard* shard);
int64_t b2pt_shard_misses(B2ptShard* shard);

/* Device pointers of the accumulators (W*H*3 floats), for zero-copy hand-off
 * to a collective or a device-side denoiser. */
float* b2pt_device_image(B2ptCtx* ctx);
float* b2pt_device_albedo(B2ptCtx* ctx);

/* Accumulate into caller-owned device memory (e.g. an NCCL-registered
 * buffer) instead of the context's own image.  NULL restores the default. */
int b2pt_set_device_image(B2ptCtx* ctx, float* image_dev);

/* The CUDA stream (cudaStream_t) all work of this context is queued on. */
void* b2pt_stream(B2ptCtx* ctx);

/* Queue all further work on a caller-owned stream (e.g. the stream a
 * collective library runs on, so render -> reduce needs no host sync).
 * NULL restores the context's own stream.  Synchronises the old stream. */
int b2pt_set_stream(B2ptCtx* ctx, void* cuda_stream);

/* Mirror of timer().getGpuElapsedTimeForPreviousOperation()
 * (apps/src/main.cpp:263): milliseconds of the depth loop of the last
 * rendered iteration (the window of apps/src/pathtrace.cu:583-653). */
float b2pt_last_loop_ms(B2ptCtx* ctx);

/* sendImageToPBO / sendDenosiedImageToPBO (apps/src/pathtrace.cu:73-116):
 * rgba8_dev[i] = clamp(int(pix / iter * 255.0)), w = 0.  `src_dev` NULL means
 * the accumulation image; iter <= 0 means "already normalised" (the denoised
 * variant).  rgba8_dev is device memory, W*H*4 bytes. */
int b2pt_tonemap_rgba8(B2ptCtx* ctx, const float* src_dev, int32_t iter, uint8_t* rgba8_dev);

/* ---- output hand-off (saveImage, CPUdenoise) ---------------------------------- */
enum { B2PT_AOV_IMAGE = 0, B2PT_AOV_ALBEDO = 1 };

/* The pixels saveImage() hands to image::savePNG (apps/src/main.cpp:115-135,
 * apps/src/image.cpp:22-33), quantised on the device:
 *     v = image / samples (B2PT_AOV_IMAGE)  or  albedo (B2PT_AOV_ALBEDO)
 *     rgb8[(y*W + x')*3 + c] = (unsigned char)(clamp(v, 0, 1) * 255.f)
 * with x' = W-1-x when mirror_x != 0 (saveImage mirrors, main.cpp:126).
 * rgb8_host holds W*H*3 bytes.  Synchronous. */
int b2pt_resolve_rgb8(B2ptCtx* ctx, int32_t aov, int32_t samples, int32_t mirror_x, uint8_t* rgb8_host);

/* saveImage + image::savePNG: write the mirrored, quantised AOV as an 8-bit RGB
 * PNG at `path` (the caller composes the "<FILE>.<time>.<n>samp.png" name). */
int b2pt_save_png(B2ptCtx* ctx, int32_t aov, int32_t samples, const char* path);

/* saveImage + image::saveHDR (apps/src/image.cpp:41-45; the call is commented
 * out at apps/src/main.cpp:163): the same mirrored AOV, unquantised, as a
 * Radiance .hdr file.  The RGBE pixels are the ones stbi_write_hdr produces;
 * scanlines are stored flat instead of run-length encoded. */
int b2pt_save_hdr(B2ptCtx* ctx, int32_t aov, int32_t samples, const char* path);

/* The denoiser's "color" input (CPUdenoise, apps/src/main.cpp:189-203):
 * color = image / (float)iter as W*H Float3, written to color_dev (device
 * memory, may be NULL) and/or color_host (may be NULL).  Together with
 * b2pt_device_albedo() these are the two images oidn::Filter::setImage takes,
 * so a device-side denoiser needs no host round trip. */
int b2pt_resolve_color(B2ptCtx* ctx, int32_t iter, float* color_dev, float* color_host);

/* ---- statistics and stage dumps (parity / measurement) --------------------- */

/* Live path counts of the last rendered iteration: n_live[d] = number of
 * paths entering depth d (n_live[0] = W*H).  Returns the number of depths
 * written (<= cap).  Synchronises. */
int b2pt_live_counts(B2ptCtx* ctx, int32_t* n_live, int32_t cap);

/* Mesh-walk queue lengths of the last rendered iteration: walks[d] = rays of
 * depth d that had to walk a mesh BVH, long_walks[d] = (ray, mesh) walks handed
 * to the warp-cooperative kernel (may be NULL).  Returns the depths written. */
int b2pt_walk_counts(B2ptCtx* ctx, int32_t* walks, int32_t* long_walks, int32_t cap);

/* Per-kernel device time of ONE iteration, measured with CUDA events around
 * every launch (no graph): ms[0] generate, ms[1] intersect (sum over depths),
 * ms[2] material sort, ms[3] shade+compact+gather, ms[4] whole iteration.
 * The iteration is rendered and accumulated like any other.  Synchronous. */
int b2pt_profile_iteration(B2ptCtx* ctx, int32_t iter, float ms[5]);

/* The same measurement per kernel: the intersect stage is four kernels
 * (analytic geoms; per-lane BVH walk; warp-cooperative long walks; record
 * finishing).  Sums over the depths of one iteration, milliseconds. */
enum {
  B2PT_PROF_GENERATE = 0,  /* k_generate                                      */
  B2PT_PROF_ANALYTIC = 1,  /* k_intersect_analytic                            */
  B2PT_PROF_WALK = 2,      /* k_mesh_walk                                     */
  B2PT_PROF_WALK_LONG = 3, /* k_mesh_walk_long                                */
  B2PT_PROF_FINISH = 4,    /* k_mesh_finish                                   */
  B2PT_PROF_SORT = 5,      /* k_sort_material                                 */
  B2PT_PROF_SHADE = 6,     /* k_shade_compact                                 */
  B2PT_PROF_ITERATION = 7, /* the whole iteration                             */
  B2PT_PROF_COUNT = 8
};
int b2pt_profile_kernels(B2ptCtx* ctx, int32_t iter, float ms[B2PT_PROF_COUNT]);

/* Kernel launches issued by this context since creation. */
int64_t b2pt_launch_count(B2ptCtx* ctx);

/* Stage ids for b2pt_stage_read; all arrays are in slot order. */
enum {
  B2PT_STAGE_RAY_ORIGIN = 0, /* float[n*3]  rays entering depth d (pre-sort)  */
  B2PT_STAGE_RAY_DIR = 1,    /* float[n*3]                                    */
  B2PT_STAGE_RAY_PIXEL = 2,  /* int32[n]    pixelIndex, pre-sort order        */
  B2PT_STAGE_HIT_T = 3,      /* float[n]    -1 on a miss, pre-sort order      */
  B2PT_STAGE_HIT_NORMAL = 4, /* float[n*3]                                    */
  B2PT_STAGE_HIT_UV = 5,     /* float[n*2]  valid for OBJ hits only           */
  B2PT_STAGE_HIT_GEOM = 6,   /* int32[n]    -1 on a miss                      */
  B2PT_STAGE_HIT_FACE = 7,   /* int32[n]    face within the geom, -1 if none  */
  B2PT_STAGE_HIT_MATERIAL = 8, /* int32[n]  0 on a miss (the memset value)    */
  B2PT_STAGE_SORT_PERM = 9,  /* int32[n]    sorted slot j <- pre-sort slot    */
  B2PT_STAGE_SHADED_COLOR = 10, /* float[n*3] after shade, sorted order       */
  B2PT_STAGE_SHADED_BOUNCES = 11, /* int32[n] remainingBounces after shade    */
  B2PT_STAGE_SHADED_ORIGIN = 12,  /* float[n*3]                               */
  B2PT_STAGE_SHADED_DIR = 13,     /* float[n*3]                               */
  B2PT_STAGE_PARTITION_PIXEL = 14 /* int32[n] pixelIndex after the stable
                                     partition: live prefix then dead tail   */
};

/* Copy one stage array of depth `depth` of the last iteration rendered with
 * record_stages=1 into host memory.  `bytes` is the capacity of `dst`.
 * Returns the number of bytes written or a negative error. */
int64_t b2pt_stage_read(B2ptCtx* ctx, int32_t depth, int32_t stage, void* dst, int64_t bytes);

/* ---- BVH introspection ------------------------------------------------------ */
typedef struct B2ptBvhInfo {
  int32_t n_faces;
  int32_t n_nodes;
  int32_t max_depth;
  float build_ms; /* Morton + radix sort + hierarchy + refit, device time */
} B2ptBvhInfo;
int b2pt_bvh_info(B2ptCtx* ctx, int32_t geom, B2ptBvhInfo* info);

/* ---- standalone device primitives -------------------------------------------
 * The scan / compaction / sort kernels of the wavefront loop, callable on
 * host arrays.  They serve the surface of the reference's side library
 * StreamCompaction::{Efficient,Naive,Thrust}::scan / compact
 * (apps/stream_compaction/efficient.h:8-11) and make the kernels unit-testable.
 */

/* Exclusive prefix sum of n ints (decoupled look-back single pass). */
int b2pt_scan_exclusive_i32(int32_t n, int32_t* out_host, const int32_t* in_host);

/* Stable compaction of the non-zero elements; returns the count or <0. */
int b2pt_compact_nonzero_i32(int32_t n, int32_t* out_host, const int32_t* in_host);

/* Stable partition permutation on a predicate array: perm_host[k] is the
 * source index of output slot k (flag!=0 first, both halves stable), the
 * permutation thrust::stable_partition applies at pathtrace.cu:649.
 * Returns the number of kept elements or <0. */
int b2pt_partition_perm(int32_t n, const uint8_t* flags_host, int32_t* perm_host);

/* Stable sort permutation by DESCENDING key (pathtrace.cu:512-516,612):
 * perm_host[k] = source index of sorted slot k.  keys must be in [0, 65535]. */
int b2pt_sort_desc_perm(int32_t n, const int32_t* keys_host, int32_t* perm_host);

/* The renderer's material sort with its compaction ranks, on host arrays (the kernel behind
 * thrust::sort_by_key(sortByMaterial) + the live prefix of thrust::stable_partition,
 * pathtrace.cu:512-516,612,649): material_host[i] < n_materials is the material of path slot i,
 * live_host[i] != 0 says the path survives the shade.  perm_host[j] = slot of sorted position j
 * (stable, descending material); rank_host[j] = survivors in front of sorted position j.
 * general != 0 forces the 256-bin kernel where the few-materials kernel (n_materials <= 8) applies.
 * Returns the number of survivors or <0. */
int b2pt_sort_material_ranks(int32_t n, const uint8_t* material_host, const uint8_t* live_host, int32_t n_materials,
                             int32_t general, int32_t* perm_host, int32_t* rank_host);

/* Stable LSD radix sort of 32-bit keys with 32-bit values (onesweep, 4 passes
 * of 8 bits); the sort behind the LBVH Morton ordering. */
int b2pt_radix_sort_pairs_u32(int32_t n, uint32_t* keys_host, uint32_t* vals_host);

/* ---- scene loader -------------------------------------------------------------
 * The scenes/<name>.txt format of apps/src/scene.cpp (MATERIAL / CAMERA / OBJECT
 * blocks), OBJ+MTL meshes and 8-bit textures.  The camera returned is the one
 * the renderer uses, i.e. after the orbit recompute of main.cpp:67-81,222-240.
 */
typedef struct B2ptLoadedScene B2ptLoadedScene;

typedef struct B2ptLoadOverrides {
  int32_t width;      /* >0 overrides RES x (pixelLength is recomputed)      */
  int32_t height;     /* >0 overrides RES y                                  */
  int32_t iterations; /* >0 overrides ITERATIONS                             */
  int32_t depth;      /* >0 overrides DEPTH                                  */
  int32_t per_face_materials; /* 1: keep tinyobj's per-face material ids: every
                         MTL material of an OBJ becomes a scene material with its
                         own four maps (B2ptScene::face_material /
                         material_textures); 0 (def): the reference's one material
                         per OBJ                                               */
} B2ptLoadOverrides;

int b2pt_scene_load(const char* path, const B2ptLoadOverrides* ov, B2ptLoadedScene** out);
const B2ptScene* b2pt_scene_view(const B2ptLoadedScene* s);
const char* b2pt_scene_image_name(const B2ptLoadedScene* s); /* FILE line */
/* What the loader tolerated the way the reference does instead of failing, one line each ("" if nothing): today the
 * texture files an MTL names that are missing or cannot be decoded -- the reference prints "Failed to load ... texture
 * file" and renders without the map (scene.cpp:150-154), and so does this library. */
const char* b2pt_scene_warnings(const B2ptLoadedScene* s);
void b2pt_scene_free(B2ptLoadedScene* s);

/* ---- misc ------------------------------------------------------------------------- */
const char* b2pt_last_error(void);
int b2pt_abi_version(void);
/* Number of visible CUDA devices, or a negative error. */
int b2pt_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B2PT_H_ */
