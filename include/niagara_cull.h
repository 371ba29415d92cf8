/*
 * niagara_cull.h — C ABI of the B200-native visibility path for zeux/niagara.
 *
 * This library replaces the Vulkan compute dispatches of niagara's GPU-driven visibility path
 * (reference call sites, all in src/niagara.cpp):
 *
 *   cull lambda      niagara.cpp:1530-1574   drawcull.comp (+ tasksubmit.comp)    -> nvc_drawcull
 *   render lambda    niagara.cpp:1582-1610   clustercull.comp + clustersubmit.comp -> nvc_clustercull
 *   pyramid lambda   niagara.cpp:1703-1733   depthreduce.comp, one dispatch / mip  -> nvc_depth_pyramid
 *   task-shader path niagara.cpp:1666-1679   meshlet.task.glsl (payload variant)   -> nvc_taskcull
 *
 * with hand-written sm_100a CUDA kernels.  The buffer contract (struct layouts, counters, padding,
 * overflow behaviour) is the reference's, bit for bit, so the outputs can be consumed by the existing
 * indirect-draw path (vkCmdDrawIndexedIndirectCount / vkCmdDrawMeshTasksIndirectEXT) unchanged.
 *
 * Conventions
 *   - plain C, no CUDA or torch types in signatures: streams are passed as void* (a cudaStream_t),
 *     device buffers as plain pointers to the structs below.
 *   - every pass call only ENQUEUES work on `stream`; no allocation, no synchronisation, safe to capture
 *     into a CUDA graph.  A context is not re-entrant (the reference's frame loop is single threaded): its passes
 *     share device-side ticket counters, so the passes of ONE context must be ordered — same stream, or streams
 *     ordered by events; use one context per concurrently running frame.  The context's device must be the calling
 *     thread's current CUDA device for pass calls (else NVC_ERROR_INVALID_ARGUMENT); nvc_create makes `device` current
 *     and leaves it so, the other lifetime calls restore the caller's device.
 *   - return value: 0 = NVC_OK, negative = NvcStatus error.  Capacity overflow (TASK_WGLIMIT /
 *     CLUSTER_LIMIT) is NOT an error: it is the reference's defined drop-and-clamp behaviour.
 *   - all struct layouts below are asserted (size/offset) against src/shaders/mesh.h + src/scene.h.
 */
#ifndef NIAGARA_CULL_H
#define NIAGARA_CULL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NVC_API __declspec(dllexport)
#else
#define NVC_API __attribute__((visibility("default")))
#endif

/* ---- limits: src/config.h:1-28 ------------------------------------------------------------ */
#define NVC_TASK_WGSIZE 64u           /* config.h:2  meshlets per task command / cluster-cull group  */
#define NVC_TASK_WGLIMIT (1u << 22)   /* config.h:25 max task commands                               */
#define NVC_CLUSTER_LIMIT (1u << 24)  /* config.h:28 max visible cluster indices                     */
#define NVC_CLUSTER_TILE 16u          /* config.h:22 X dimension of the 16 x Y x 16 mesh dispatch    */
#define NVC_MAX_DISPATCH_GROUPS 65535u /* tasksubmit.comp.glsl:36, clustersubmit.comp.glsl:37       */
#define NVC_MAX_LODS 8u               /* mesh.h:77   Mesh::lods[8]                                   */
#define NVC_MAX_HIZ_LEVELS 16u        /* sampler maxLod 16 (resources.cpp:306)                       */

/* ---- data layouts: src/shaders/mesh.h:11-123 == src/scene.h:10-93 == src/niagara.cpp:227-260 ---- */

/* mesh.h:11-24 / scene.h:10-23 — 24 B, 8-byte aligned.  Cull reads bytes 0..11 only. */
typedef struct NvcMeshlet
{
	uint16_t center[3]; /* IEEE binary16 */
	uint16_t radius;    /* IEEE binary16 */
	int8_t cone_axis[3];
	int8_t cone_cutoff;
	uint32_t dataOffset;
	uint32_t baseVertex;
	uint8_t vertexCount;
	uint8_t triangleCount;
	uint8_t shortRefs;
	uint8_t padding;
} NvcMeshlet;

/* mesh.h:53-60 / scene.h:68-75 — 20 B */
typedef struct NvcMeshLod
{
	uint32_t indexOffset;
	uint32_t indexCount;
	uint32_t meshletOffset;
	uint32_t meshletCount;
	float error;
} NvcMeshLod;

/* mesh.h:62-78 / scene.h:77-93 — 208 B, 16-byte aligned */
typedef struct NvcMesh
{
	float center[3];
	float radius;
	uint32_t vertexOffset;
	uint32_t vertexCount;
	uint32_t ommIndexData;
	uint32_t ommIndexBase;
	uint32_t lodCount;
	uint32_t lodRT;
	uint32_t padding[2];
	NvcMeshLod lods[NVC_MAX_LODS];
} NvcMesh;

/* mesh.h:92-102 / scene.h:39-49 — 48 B, 16-byte aligned; orientation stored x,y,z,w */
typedef struct NvcMeshDraw
{
	float position[3];
	float scale;
	float orientation[4];
	uint32_t meshIndex;
	uint32_t meshletVisibilityOffset;
	uint32_t postPass;
	uint32_t materialIndex;
} NvcMeshDraw;

/* mesh.h:104-114 / niagara.cpp:227-231 — 24 B: drawId + VkDrawIndexedIndirectCommand */
typedef struct NvcMeshDrawCommand
{
	uint32_t drawId;
	uint32_t indexCount;
	uint32_t instanceCount;
	uint32_t firstIndex;
	uint32_t vertexOffset;
	uint32_t firstInstance;
} NvcMeshDrawCommand;

/* mesh.h:116-123 / niagara.cpp:233-240 — 20 B */
typedef struct NvcMeshTaskCommand
{
	uint32_t drawId;
	uint32_t taskOffset;
	uint32_t taskCount;
	uint32_t lateDrawVisibility;
	uint32_t meshletVisibilityOffset;
} NvcMeshTaskCommand;

/* mesh.h:26-44 / niagara.cpp:242-260 — 144 B (alignas 16), passed verbatim (the reference's push constants) */
typedef struct NvcCullData
{
	float view[16]; /* column-major mat4 */
	float P00, P11, znear, zfar;
	float frustum[4];
	float lodTarget;
	float pyramidWidth, pyramidHeight;
	uint32_t drawCount;
	int32_t cullingEnabled;
	int32_t lodEnabled;
	int32_t occlusionEnabled;
	int32_t clusterOcclusionEnabled;
	int32_t clusterBackfaceEnabled;
	uint32_t postPass;
	uint32_t pad_[2];
} NvcCullData;

/* mesh.h:125-128 — task-shader payload (meshlet.task.glsl:47) */
typedef struct NvcMeshTaskPayload
{
	uint32_t clusterIndices[NVC_TASK_WGSIZE];
} NvcMeshTaskPayload;

/* Depth pyramid ("depthPyramid" image, niagara.cpp:1339-1350): R32F, width x height = previousPow2 of the
 * depth target, full mip chain.  In place of a VkImage: one linear device allocation, level l stored
 * row-major and tightly packed at texels + level_offset[l], size max(1,width>>l) x max(1,height>>l). */
typedef struct NvcHiZ
{
	float* texels; /* device pointer */
	uint32_t width, height, levels;
	uint32_t level_offset[NVC_MAX_HIZ_LEVELS]; /* in texels */
	uint32_t total_texels;
} NvcHiZ;

typedef struct NvcLimits
{
	uint32_t task_wglimit;  /* default NVC_TASK_WGLIMIT  */
	uint32_t cluster_limit; /* default NVC_CLUSTER_LIMIT */
} NvcLimits;

typedef enum NvcStatus
{
	NVC_OK = 0,
	NVC_ERROR_INVALID_ARGUMENT = -1,
	NVC_ERROR_CUDA = -2,
	NVC_ERROR_NO_DEVICE = -3,
	NVC_ERROR_OUT_OF_MEMORY = -4,
	NVC_ERROR_NCCL = -5,
	NVC_ERROR_UNSUPPORTED = -6, /* valid input this build has no decoder for (see nvc_scene_cache_read) */
	NVC_ERROR_CORRUPT = -7      /* a scene cache whose sections do not add up */
} NvcStatus;

typedef struct NvcContext NvcContext;

/* ---- lifetime -------------------------------------------------------------------------------- */

/* Replaces pipeline/program creation for the cull path (niagara.cpp:652-762).  `limits` may be NULL. */
NVC_API int nvc_create(int device, const NvcLimits* limits, NvcContext** out_ctx);
NVC_API void nvc_destroy(NvcContext* ctx);
NVC_API const char* nvc_status_string(int status);
/* last CUDA error text recorded by this context (for NVC_ERROR_CUDA) */
NVC_API const char* nvc_last_error(const NvcContext* ctx);
NVC_API const char* nvc_version(void);

/* Optional, once per geometry upload (where the reference uploads `mb`, niagara.cpp:1040-1046): builds a context-owned,
 * read-only cull view of Mesh[] — a 32-byte head per mesh {center, radius, lodCount, lods[0].meshletOffset/Count,
 * vertexOffset} plus the 8 LOD errors — so that nvc_drawcull touches ONE 32-byte sector per draw for single-LOD / LOD-0
 * task draws instead of the 3-4 sectors the 208-byte AoS record spreads those fields over.  Used automatically by
 * nvc_drawcull calls that pass the same `meshes` pointer; results are identical.  The Mesh[] contents must not change
 * afterwards (call again after re-uploading; meshes = NULL releases the view). */
NVC_API int nvc_prepare_meshes(NvcContext* ctx, void* stream, const NvcMesh* meshes, uint32_t mesh_count);

/* Tuning: stage the coarse tail of the depth pyramid (top mips, at most `texels` texels, capped at 11264 = 44 KB) into
 * shared memory once per CTA of the late cluster pass with one TMA bulk copy (cp.async.bulk, SASS UBLKCP); lookups
 * whose whole warp samples a staged mip then read shared memory.  0 (default) disables it: on the measured workloads
 * the shared-memory carve-out costs more L1 hit rate on the fine mips than the staged mips save.
 * Results are identical either way.  Also settable with the environment variable NVC_HIZ_STAGE_TEXELS. */
NVC_API int nvc_set_hiz_staging(NvcContext* ctx, uint32_t texels);

/* Derived, optional, results identical (like nvc_prepare_meshes): allocates a context-owned FOOTPRINT IMAGE for this pyramid
 * (same texel count + one row / column per mip).  nvc_depth_pyramid then also writes, for every mip and every 2 x 2 sampler
 * footprint, the minimum of its four (edge-clamped) texels; the late cluster pass reads ONE value per meshlet instead of four
 * scattered texels.  The image is valid only for pyramids written by nvc_depth_pyramid: a caller that stores texels itself must
 * not prepare (or must pass NULL here to release).  Allocates / frees: call outside stream capture.  hiz == NULL releases. */
NVC_API int nvc_prepare_hiz(NvcContext* ctx, const NvcHiZ* hiz);
/* Diagnostics / tests: the device pointer of the footprint image, its first mip (mips below it have no image), the entry offset of
 * every mip (offset_out[NVC_MAX_HIZ_LEVELS], floats) and the total entry count.  Mip l of width w, height h: rows of pitch
 * (w + 4) & ~3 entries, entry (i + 1, j + 1) = min of texels (i..i+1, j..j+1) clamped to the mip, i in [-1, w-1], j in [-1, h-1].
 * Valid after the nvc_depth_pyramid call that wrote it; *image_out = NULL when no image is prepared. */
NVC_API int nvc_hiz_footprints(NvcContext* ctx, const float** image_out, uint32_t* first_level_out, uint32_t* offset_out, uint32_t* total_out);

/* The cluster pass runs by default as a conservative FILTER (fused arithmetic with an error margin on every comparison,
 * per-command transforms) whose undecided meshlets are re-evaluated by the exact strict-IEEE path: results are identical,
 * the pass is ~2x faster.  enabled = 0 selects the exact kernel for every meshlet (A/B measurements, debugging). */
NVC_API int nvc_set_cluster_filter(NvcContext* ctx, int enabled);
/* Diagnostics (synchronises the device): out[0] = meshlets the filtered cluster passes evaluated, out[1] = how many of them
 * were undecided and took the exact path, both cumulative since the last call with reset != 0. */
NVC_API int nvc_filter_stats(NvcContext* ctx, uint64_t* out_items_undecided2, int reset);

/* ---- pyramid layout (host only): niagara.cpp:439-447 previousPow2, resources.cpp:280-292 getImageMipLevels,
 *      niagara.cpp:1339-1342 ------------------------------------------------------------------------- */
NVC_API uint32_t nvc_previous_pow2(uint32_t v);
NVC_API uint32_t nvc_image_mip_levels(uint32_t width, uint32_t height);
/* Fills width/height/levels/level_offset/total_texels for a depth target of depth_width x depth_height;
 * leaves `texels` NULL (caller allocates total_texels * 4 bytes on the device). */
NVC_API int nvc_hiz_layout(uint32_t depth_width, uint32_t depth_height, NvcHiZ* out);

/* ---- passes ---------------------------------------------------------------------------------- */

/* drawcull.comp.glsl:54-156 for all four (LATE, TASK) specialisations; when task != 0 the
 * tasksubmit.comp.glsl:27-47 epilogue (group counts + zero padding to x64) is fused into the same launch.
 *   draws            MeshDraw[cull->drawCount]                      (db)
 *   meshes           Mesh[]                                         (mb)
 *   draw_visibility  uint32[drawCount], read; written when late     (dvb)
 *   commands         task ? MeshTaskCommand[task_wglimit] : MeshDrawCommand[drawCount]   (dcb)
 *   command_count4   uint32[4] {commandCount, groupCountX, groupCountY, groupCountZ}      (dccb)
 *                    fully written by the call (the host-side vkCmdFillBuffer reset of niagara.cpp:1541 is
 *                    folded in); for task == 0 only [0] is meaningful, [1..3] are set to 0.
 *   hiz              depth pyramid; required when late && cull->occlusionEnabled == 1, else may be NULL */
NVC_API int nvc_drawcull(NvcContext* ctx, void* stream, const NvcCullData* cull, int late, int task,
    const NvcMeshDraw* draws, const NvcMesh* meshes, uint32_t* draw_visibility,
    void* commands, uint32_t* command_count4, const NvcHiZ* hiz);

/* clustercull.comp.glsl:56-149 (LATE = 0/1) with the clustersubmit.comp.glsl:25-45 epilogue fused.
 * Processes commandId < command_count4[1] * 64 exactly as the reference's indirect dispatch (X,64,1) does.
 *   task_commands       MeshTaskCommand[] written by nvc_drawcull(task=1)   (dcb)
 *   command_count4      uint32[4] written by nvc_drawcull(task=1)           (dccb)
 *   meshlets            Meshlet[]                                            (mlb)
 *   meshlet_visibility  bit array, read; atomically updated when late        (mvb)
 *   cluster_indices     uint32[cluster_limit rounded up to 256]              (cib)
 *   cluster_count4      uint32[4] {clusterCount, 16, Y, 16}, fully written   (ccb) */
NVC_API int nvc_clustercull(NvcContext* ctx, void* stream, const NvcCullData* cull, int late,
    const NvcMeshTaskCommand* task_commands, const uint32_t* command_count4,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, uint32_t* meshlet_visibility,
    uint32_t* cluster_indices, uint32_t* cluster_count4, const NvcHiZ* hiz);

/* meshlet.task.glsl:53-149 (task-shader submission mode, niagara.cpp:1666-1679): same per-meshlet test as
 * nvc_clustercull, but the survivors of command c are compacted into payloads[c].clusterIndices[0..n) and
 * n is written to emit_counts[c] (the EmitMeshTasksEXT(n,1,1) argument).  Order inside a payload is
 * ascending lane (the reference's shared-memory atomic makes it arbitrary). */
NVC_API int nvc_taskcull(NvcContext* ctx, void* stream, const NvcCullData* cull, int late,
    const NvcMeshTaskCommand* task_commands, const uint32_t* command_count4,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, uint32_t* meshlet_visibility,
    NvcMeshTaskPayload* payloads, uint32_t* emit_counts, const NvcHiZ* hiz);

/* Consumer-side walk of the cluster pass output, exactly as meshlet.mesh.glsl:89-105 decodes it under
 * vkCmdDrawMeshTasksIndirectEXT(ccb, 4): one slot per mesh workgroup, slot = x + y * 256 + z * 16 for the dispatch
 * (16, Y, 16) stored in cluster_count4[1..3]; ci == ~0 -> skipped; command = taskCommands[ci & 0xffffff];
 * mi = command.taskOffset + (ci >> 24).  Writes per slot {drawId, mi, vertexCount, triangleCount} (all ~0 for skipped
 * slots) and stats[4] = {decoded slots, skipped slots, invalid slots (lane >= taskCount), triangles}.  Used to prove the
 * output is a drop-in for the reference's consumers; not part of the per-frame path.  records may be NULL (stats only). */
typedef struct NvcClusterRecord
{
	uint32_t drawId, meshletIndex, vertexCount, triangleCount;
} NvcClusterRecord;
NVC_API int nvc_decode_clusters(NvcContext* ctx, void* stream, const uint32_t* cluster_indices, const uint32_t* cluster_count4,
    const NvcMeshTaskCommand* task_commands, const NvcMeshlet* meshlets, NvcClusterRecord* records, uint32_t* stats4);

/* depthreduce.comp.glsl:14-22 + the per-mip loop of niagara.cpp:1703-1733, as ONE launch.
 *   depth  float[depth_height][depth_width], reverse-Z (the D32 depthTarget)
 *   hiz    layout from nvc_hiz_layout(depth_width, depth_height) with `texels` set */
NVC_API int nvc_depth_pyramid(NvcContext* ctx, void* stream, const float* depth,
    uint32_t depth_width, uint32_t depth_height, const NvcHiZ* hiz);

/* ---- host-side helpers mirroring the reference's host code (pure CPU, no context needed) ---------- */

/* niagara.cpp:449-481 PCG32 + rand01/rand32 and the random scene of niagara.cpp:969-998.
 * Fills draws[draw_count] exactly as the reference does (rng state = 0x42, inc = PCG32 default). */
NVC_API void nvc_host_random_draws(NvcMeshDraw* draws, uint32_t draw_count, uint32_t mesh_count, float scene_radius);

/* niagara.cpp:1002-1020: meshletVisibilityOffset = exclusive prefix sum over draws of max-over-LODs
 * meshletCount; returns the total bit count; *post_pass_mask gets OR(1 << postPass) (may be NULL). */
NVC_API uint32_t nvc_host_visibility_offsets(NvcMeshDraw* draws, uint32_t draw_count, const NvcMesh* meshes, uint32_t* post_pass_mask);

typedef struct NvcCamera
{
	float position[3];
	float orientation[4]; /* x,y,z,w */
	float fovY;
	float znear;
} NvcCamera;

typedef struct NvcCullOptions
{
	float draw_distance; /* zfar, niagara.cpp:1000 (200) */
	int32_t culling, lod, occlusion, cluster_occlusion, mesh_shading; /* runtime toggles niagara.cpp:31-44 */
	int32_t debug_lod_step;
} NvcCullOptions;

/* niagara.cpp:424-432 + 1487-1516: view, infinite reverse-Z projection, frustum planes, lodTarget, pyramid size.
 * Leaves clusterBackfaceEnabled = 0 and postPass = 0 exactly like `cullData` at niagara.cpp:1499-1516;
 * the per-pass copies (niagara.cpp:1547-1550, 1595-1596) are the caller's job, see nvc_host_pass_data. */
NVC_API void nvc_host_cull_data(const NvcCamera* camera, uint32_t screen_width, uint32_t screen_height,
    uint32_t draw_count, const NvcCullOptions* options, NvcCullData* out, float* out_projection16 /* may be NULL */);

/* passData of the cull lambda (niagara.cpp:1547-1550): clusterBackfaceEnabled = (postPass == 0), postPass set. */
NVC_API void nvc_host_pass_data(const NvcCullData* frame, int for_drawcull, uint32_t post_pass, NvcCullData* out);

/* ---- multi-GPU: exchange of the per-rank visible slabs (new in this implementation; the reference is single-GPU) */

/* Bootstraps an NCCL communicator owned by the context.  unique_id is the 128-byte ncclUniqueId produced by
 * nvc_nccl_unique_id on rank 0 and distributed by the caller (e.g. torch.distributed broadcast). */
NVC_API int nvc_nccl_unique_id(void* out_unique_id128);
NVC_API int nvc_nccl_init(NvcContext* ctx, const void* unique_id128, int rank, int world_size);
/* Gathers every rank's {count, slab[0..slab_capacity)} into gathered_counts[world] and
 * gathered[world][slab_capacity] on every rank (one ncclAllGather each), after globalising nothing: ids inside
 * the slabs are rank-local; rank r's base draw id is the caller's partition offset. */
NVC_API int nvc_allgather_visible(NvcContext* ctx, void* stream, const void* local_slab, size_t slab_bytes,
    const uint32_t* local_count4, void* gathered_slabs, uint32_t* gathered_count4);

/* Copy-engine variant of the same exchange (no SMs, so it overlaps the SM-filling late cluster pass): every rank
 * PUSHES its slab + counters into slot `rank` of every rank's gathered buffers with peer cudaMemcpyAsync over CUDA-IPC
 * mappings, then raises a flag in the peer's memory; nvc_gather_wait blocks a stream (device side) until all ranks'
 * data of the latest push has landed.  One process per GPU.
 *   nvc_gather_create   allocates this rank's receive buffers (double-buffered by frame parity), returns a 192-byte IPC ticket
 *   nvc_gather_connect  all ranks' tickets (world x 192 bytes, rank order, exchanged by the caller)
 *   nvc_gather_push     enqueue after the pass that produced local_slab / local_count4 on `stream`; the counters are
 *                       snapshotted in stream order, the slab must stay unchanged until nvc_gather_wait
 *   nvc_gather_wait     enqueue on the stream that consumes the gathered data (or reuses local_slab); every push must be
 *                       followed by its wait before the second-next push
 *   nvc_gather_buffers  this rank's gathered slabs [world][slab_bytes] and counters [world][4] of the latest frame, valid from
 *                       its nvc_gather_wait until the next nvc_gather_wait
 * No barrier is needed between frames: receivers acknowledge to senders which frame they are done with, and a sender reuses a
 * parity buffer only after every peer has acknowledged its previous user (a fast rank runs at most one frame ahead). */
NVC_API int nvc_gather_create(NvcContext* ctx, size_t slab_bytes, int rank, int world_size, void* ticket192_out);
NVC_API int nvc_gather_connect(NvcContext* ctx, const void* all_tickets);
NVC_API int nvc_gather_push(NvcContext* ctx, void* stream, const void* local_slab, const uint32_t* local_count4);
NVC_API int nvc_gather_wait(NvcContext* ctx, void* stream);
NVC_API int nvc_gather_buffers(NvcContext* ctx, void** gathered_slabs, uint32_t** gathered_count4);
/* CUDA-graph replay of frames that contain the exchange: the calls above bake their frame tags (and the parity-selected buffer
 * addresses) into a captured graph.  Enqueue this as the LAST operation of the captured sequence with `frames` = the number of
 * nvc_gather_push calls captured (must be even); every replay then shifts the tags of the next one by `frames`.  Each captured push
 * needs its nvc_gather_wait inside the same capture; all ranks replay the same graphs the same number of times. */
NVC_API int nvc_gather_graph_advance(NvcContext* ctx, void* stream, uint32_t frames);
/* The device-side waits of the protocol give up after ~20 s (a peer that crashed or disagrees about the frame) instead of hanging the
 * GPU; *timed_out = 1 from then on (the gathered data are undefined, later waits return at once).  Synchronises the device. */
NVC_API int nvc_gather_status(NvcContext* ctx, int* timed_out);
/* transport of nvc_gather_push: 0 = copy engines (default), 2 / 3 = NVSwitch multicast (below), 1 = a 32-CTA kernel on a high-priority stream that writes only
 * the valid count x 20 bytes with 16-byte peer stores (also selectable with NVC_GATHER_MODE=sm) */
NVC_API int nvc_gather_set_mode(NvcContext* ctx, int mode);

/* NVSwitch MULTICAST transports (modes 2 and 3).  Every rank allocates ONE region of nvc_gather_region_bytes() with a symmetric
 * allocator that also yields a multicast mapping of it (CUDA VMM: cuMemCreate + cuMulticastCreate / AddDevice / BindMem, handles
 * exchanged between the processes; or torch.distributed._symmetric_memory, as bench.py does) and attaches it:
 *   peer_regions[p]    this process's mapping of rank p's region (p == rank: its own)
 *   multicast_region   the multicast alias (a store through it lands in EVERY rank's region), or NULL: unicast SM push only
 * The caller barriers once after every rank has attached.  mode 2: nvc_gather_push launches a kernel that stores the valid part
 * of the slab ONCE through the multicast mapping (egress 1x instead of world x).  mode 3 (fused compute + collective):
 *   nvc_gather_fuse_next_drawcull(ctx, stream);   // takes the frame tag, waits for the peers' acknowledgements on `stream`
 *   nvc_drawcull(ctx, stream, ..., late = 1, task = 1, ...);   // its command write-out also goes through the multicast mapping
 *   nvc_gather_push(ctx, stream, dcb, dccb);      // counters + flags only
 *   ...late cluster pass...;  nvc_gather_wait(ctx, stream); */
NVC_API size_t nvc_gather_region_bytes(size_t slab_bytes, int world_size);
NVC_API int nvc_gather_attach(NvcContext* ctx, size_t slab_bytes, int rank, int world_size, void* const* peer_regions, void* multicast_region);
NVC_API int nvc_gather_fuse_next_drawcull(NvcContext* ctx, void* stream);

/* ---- widening N2: scene cache (.cache v7) reader — the on-disk format that feeds the path ----------------
 * Replaces loadSceneCache (src/scenecache.cpp:273-370) for the arrays the visibility path consumes.  Host-only, no
 * CUDA calls: the caller maps the file, parses it once and copies Meshlet[] / Mesh[] / MeshDraw[] / Animation[] /
 * Keyframe[] straight from the mapping to the device (they are stored raw, scenecache.cpp:170,181-186).  The
 * meshopt-compressed sections are located by the header's byte counts and decoded by nvc_scene_cache_read: the
 * per-meshlet stream ("meshlet codec", scenecache.cpp:84-117,256-271), the vertex codec (vertices and RT positions) and
 * the index codec.  Every stream is bounds-checked (NVC_ERROR_CORRUPT), nothing is written outside dst. */

#define NVC_SCENE_CACHE_MAGIC 0x434E4353u /* 'SCNC', scenecache.cpp:12 */
#define NVC_SCENE_CACHE_VERSION 7u        /* scenecache.cpp:13 */

/* scenecache.cpp:16-55 — 160 bytes */
typedef struct NvcSceneCacheHeader
{
	uint32_t magic, version;
	uint64_t hashMeta;
	uint32_t meshletMaxVertices, meshletMaxTriangles;
	uint8_t clrtMode, compressed, pad0_[2];
	uint32_t compressedVertexBytes, compressedIndexBytes, compressedMeshletDataBytes, compressedMeshletVtx0Bytes;
	uint32_t vertexCount, indexCount, meshletCount, meshletdataCount, meshletvtx0Count, meshCount;
	uint32_t materialCount, drawCount, texturePathCount, lightCount, animationCount, keyframeCount;
	uint32_t ommArrayDataSize, ommIndexDataSize, ommDescCount, ommStates;
	NvcCamera camera;
	float sunDirection[3];
	uint32_t pad1_;
} NvcSceneCacheHeader;

/* sections in file order (saveSceneCache, scenecache.cpp:158-197) */
typedef enum NvcSceneCacheSectionId
{
	NVC_CACHE_VERTICES = 0, /* Vertex 16 B, vertex codec when compressed */
	NVC_CACHE_INDICES,      /* uint32, index codec when compressed */
	NVC_CACHE_MESHLETS,     /* Meshlet 24 B, always raw */
	NVC_CACHE_MESHLETDATA,  /* uint32 words, meshlet codec when compressed */
	NVC_CACHE_MESHLETVTX0,  /* uint16, vertex codec (stride 8) when compressed */
	NVC_CACHE_MESHES,       /* Mesh 208 B */
	NVC_CACHE_MATERIALS,    /* Material 64 B */
	NVC_CACHE_DRAWS,        /* MeshDraw 48 B */
	NVC_CACHE_LIGHTS,       /* Light 32 B */
	NVC_CACHE_ANIMATIONS,   /* Animation 24 B */
	NVC_CACHE_KEYFRAMES,    /* Keyframe 32 B */
	NVC_CACHE_OMM_DATA,     /* bytes */
	NVC_CACHE_OMM_INDICES,  /* bytes */
	NVC_CACHE_OMM_DESCS,    /* uint32 */
	NVC_CACHE_TEXTURE_PATHS, /* char[256] each */
	NVC_CACHE_SECTION_COUNT
} NvcSceneCacheSectionId;

typedef struct NvcSceneCacheSection
{
	uint64_t offset;        /* from the start of the file */
	uint64_t stored_bytes;  /* bytes occupied in the file */
	uint64_t decoded_bytes; /* count * element_size */
	uint32_t count, element_size;
	uint32_t compressed;    /* 1: stored_bytes is a meshopt stream, not the array */
	uint32_t pad_;
} NvcSceneCacheSection;

typedef struct NvcSceneCacheInfo
{
	NvcSceneCacheHeader header;
	NvcSceneCacheSection sections[NVC_CACHE_SECTION_COUNT];
} NvcSceneCacheInfo;

/* The three meshopt stream decoders on their own (the same streams appear in glTF EXT_meshopt_compression buffers):
 *   vertex codec v0 / v1   -> dst[vertex_count * vertex_size], vertex_size a multiple of 4, <= 256
 *   index codec v0 / v1    -> dst[index_count] uint32, index_count a multiple of 3 (triangles come back rotated)
 *   meshlet codec          -> vertex references (reference_size 2 or 4 bytes each) and triangle_count * 3 index bytes
 * Return NVC_OK, NVC_ERROR_CORRUPT (malformed stream; nothing is written outside the destinations),
 * NVC_ERROR_UNSUPPORTED (a newer codec version) or NVC_ERROR_INVALID_ARGUMENT. */
NVC_API int nvc_decode_vertex_stream(void* dst, uint32_t vertex_count, uint32_t vertex_size, const void* stream, size_t stream_size);
NVC_API int nvc_decode_index_stream(uint32_t* dst, uint32_t index_count, const void* stream, size_t stream_size);
NVC_API int nvc_decode_meshlet_stream(void* references, uint32_t vertex_count, uint32_t reference_size, uint8_t* triangles,
    uint32_t triangle_count, const void* stream, size_t stream_size);

/* scene.h:141-161 */
typedef struct NvcKeyframe
{
	float translation[3];
	float scale;
	float rotation[4]; /* x,y,z,w */
} NvcKeyframe;

typedef struct NvcAnimation
{
	int32_t drawIndex, lightIndex;
	float startTime, period;
	uint32_t keyframeOffset, keyframeCount;
} NvcAnimation;

/* Validates magic / version / that every section lies inside the file (and, when compressed, that the per-meshlet
 * stream's size chain adds up to compressedMeshletDataBytes), fills the section table.  The caller decides about
 * hashMeta / meshlet limits / clrtMode / ommStates (loadSceneCache's other rejections, scenecache.cpp:283-290). */
NVC_API int nvc_scene_cache_parse(const void* file, size_t file_size, NvcSceneCacheInfo* out);
/* Copies (raw sections) or decodes (compressed sections) one section into dst[decoded_bytes]. */
NVC_API int nvc_scene_cache_read(const void* file, size_t file_size, const NvcSceneCacheInfo* info, int section,
    void* dst, size_t dst_bytes);

/* ---- widening N3: animated MeshDraw updates (niagara.cpp:1362-1411, the frame loop's keyframe evaluation) --------------
 * Host: evaluates every animation with drawIndex >= 0 at animation_time exactly as the reference does (double index
 * arithmetic, glm::mix for position / scale, glm::slerp for orientation), writes the new values into draws[] (the
 * reference's `draws[animation.drawIndex]`) and appends {index, MeshDraw} to the packed update list.  Returns the
 * number of updates written (<= max_updates), or a negative NvcStatus. */
NVC_API int nvc_host_animate(const NvcAnimation* animations, uint32_t animation_count, const NvcKeyframe* keyframes,
    uint32_t keyframe_count, double animation_time, NvcMeshDraw* draws, uint32_t draw_count,
    uint32_t* update_indices, NvcMeshDraw* update_values, uint32_t max_updates);
/* Device: draws[update_indices[i]] = update_values[i] for i < count (the reference memcpy's into its host-visible
 * `db`, niagara.cpp:1394-1395; here the packed list is one H2D copy and one scatter launch).  Indices >= draw_count
 * are ignored; indices must be unique (the importer gives every animation its own draw, scene.cpp:777) — with
 * duplicates one of the values lands, unspecified which. */
NVC_API int nvc_update_draws(NvcContext* ctx, void* stream, NvcMeshDraw* draws, uint32_t draw_count,
    const uint32_t* update_indices, const NvcMeshDraw* update_values, uint32_t count);

/* ---- widening N4: meshlet bounds + normal cones on the GPU (the producer of Meshlet[]'s cull fields) --------------------
 * scene.h:51-57 / mesh.h:3-9 — 16 B */
typedef struct NvcVertex
{
	uint16_t vx, vy, vz; /* fp16 position */
	uint16_t tp;
	uint32_t np;
	uint16_t tu, tv;
} NvcVertex;

/* ---- N3: glTF 2.0 scene import for the visibility path (loadScene, src/scene.cpp:473-853) -----------------------------------------
 * MeshDraw[] (world transform of every mesh node decomposed as scene.cpp:295-340 does, meshIndex = first_mesh_index + running index of
 * the indexed triangle primitives, materialIndex = material_offset + material, postPass 1 for non-opaque / 2 for transmissive materials),
 * Animation[] + Keyframe[] (LINEAR TRS samplers baked per key through the node hierarchy, scene.cpp:713-830), camera, sun direction,
 * and per primitive the quantised Vertex[] / index arrays of loadVertices (scene.cpp:342-405) — bit-identical to what the reference's
 * importer produces from the same file.  Host only; .gltf (data: URIs or external buffers under base_dir) and .glb.  Sparse accessors
 * and EXT_meshopt_compression: NVC_ERROR_UNSUPPORTED.  Malformed files: NVC_ERROR_CORRUPT, never read outside the buffers. */
typedef struct NvcGltfScene NvcGltfScene;
typedef struct NvcGltfInfo
{
	uint32_t node_count, mesh_count, primitive_count, draw_count, animation_count, keyframe_count, material_count, point_light_count;
	uint32_t has_camera, has_sun;
	NvcCamera camera;        /* position, orientation, fovY of the (last) camera node; znear stays 0: the application sets it */
	float sun_direction[3];  /* local +Z of the (last) directional light node */
} NvcGltfInfo;
NVC_API int nvc_gltf_import(const void* file, size_t file_size, const char* base_dir, uint32_t first_mesh_index, uint32_t material_offset, NvcGltfScene** out_scene);
NVC_API void nvc_gltf_free(NvcGltfScene* scene);
NVC_API int nvc_gltf_info(const NvcGltfScene* scene, NvcGltfInfo* out);
/* any output may be NULL; sizes from nvc_gltf_info; mesh_scale[primitive_count] = the scale appendMesh receives (scene.cpp:503-519) */
NVC_API int nvc_gltf_scene_arrays(const NvcGltfScene* scene, NvcMeshDraw* draws, NvcAnimation* animations, NvcKeyframe* keyframes, float* mesh_scale);
NVC_API int nvc_gltf_primitive_size(const NvcGltfScene* scene, uint32_t primitive, uint32_t* vertex_count, uint32_t* index_count);
NVC_API int nvc_gltf_primitive_data(const NvcGltfScene* scene, uint32_t primitive, NvcVertex* vertices, uint32_t* indices);

/* Depth-only consumer of cib / ccb / dcb on the device (new; stands in for what the reference's mesh stage and fixed-function
 * rasteriser do to depthTarget between the cull passes, niagara.cpp:1576-1701 with meshlet.mesh.glsl:89-206): every cluster slot of
 * the (16, Y, 16) dispatch is decoded like the mesh shader does, its vertices are transformed with the shader's arithmetic
 * (strict IEEE, GLSL operation order) by `projection16 * (pass->view * vec4(world, 1))`, and its front-facing triangles are sampled at
 * pixel centres into `depth` (float, width x height, reverse Z, test GREATER: depth = max).  The caller clears `depth` to 0 before
 * the early pass; the late pass adds to it.  Meshlets whose references fall outside meshletdata / vertices are skipped and counted.
 * stats4 (may be NULL): clusters drawn, triangles sampled, clusters rejected, depth updates issued.  The image is deterministic
 * (max is order independent) and equals the sequential test rasteriser's (oracle/refshader) bit for bit. */
NVC_API int nvc_raster_depth(NvcContext* ctx, void* stream, const float* projection16, const NvcCullData* pass,
    const uint32_t* cluster_indices, const uint32_t* cluster_count4, const NvcMeshTaskCommand* task_commands,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, const uint32_t* meshletdata, uint32_t meshletdata_words,
    const NvcVertex* vertices, uint32_t vertex_count, float* depth, uint32_t width, uint32_t height, uint32_t* stats4);

/* Recomputes center / radius / cone_axis / cone_cutoff of meshlets[0..meshlet_count) from geometry, bit-identical to
 * what the reference's cooker stores (scene.cpp:69-80 appendMeshlet -> meshopt_computeMeshletBounds,
 * meshletutils.cpp:133-272,314-339): one thread per meshlet reads its references and triangles from
 * meshletdata[dataOffset...] (layout of scene.cpp:36-50, triangle order as cooked — NOT as re-read from a compressed
 * cache, whose codec rotates triangles) and the fp16 positions of vertices[baseVertex + ref].  The other Meshlet
 * fields are inputs and stay untouched.  For streamed / re-skinned geometry whose bounds must follow the vertices.
 * A meshlet whose words, references or triangle indices fall outside the given arrays is left untouched and counted
 * in *rejected (device uint32, zeroed by the call; may be NULL). */
NVC_API int nvc_cook_meshlet_bounds(NvcContext* ctx, void* stream, const NvcVertex* vertices, uint32_t vertex_count,
    const uint32_t* meshletdata, uint32_t meshletdata_words, NvcMeshlet* meshlets, uint32_t meshlet_count, uint32_t* rejected);

#ifdef __cplusplus
} /* extern "C" */

static_assert(sizeof(NvcVertex) == 16, "Vertex must match src/scene.h:51-57");
static_assert(sizeof(NvcMeshlet) == 24, "Meshlet must match src/shaders/mesh.h:11-24");
static_assert(sizeof(NvcMeshLod) == 20, "MeshLod must match src/shaders/mesh.h:53-60");
static_assert(sizeof(NvcMesh) == 208, "Mesh must match src/shaders/mesh.h:62-78");
static_assert(sizeof(NvcMeshDraw) == 48, "MeshDraw must match src/shaders/mesh.h:92-102");
static_assert(sizeof(NvcMeshDrawCommand) == 24, "MeshDrawCommand must match src/shaders/mesh.h:104-114");
static_assert(sizeof(NvcMeshTaskCommand) == 20, "MeshTaskCommand must match src/shaders/mesh.h:116-123");
static_assert(sizeof(NvcCullData) == 144, "CullData must match src/niagara.cpp:242-260 (alignas 16)");
static_assert(offsetof(NvcMesh, lods) == 48, "Mesh::lods offset");
static_assert(offsetof(NvcMesh, lodCount) == 32, "Mesh::lodCount offset");
static_assert(offsetof(NvcMeshDraw, meshIndex) == 32, "MeshDraw::meshIndex offset");
static_assert(offsetof(NvcMeshlet, cone_axis) == 8, "Meshlet::cone_axis offset");
static_assert(offsetof(NvcMeshlet, dataOffset) == 12, "Meshlet::dataOffset offset");
static_assert(offsetof(NvcCullData, P00) == 64, "CullData::P00 offset");
static_assert(offsetof(NvcCullData, frustum) == 80, "CullData::frustum offset");
static_assert(offsetof(NvcCullData, lodTarget) == 96, "CullData::lodTarget offset");
static_assert(offsetof(NvcCullData, drawCount) == 108, "CullData::drawCount offset");
static_assert(offsetof(NvcCullData, clusterBackfaceEnabled) == 128, "CullData::clusterBackfaceEnabled offset");
static_assert(offsetof(NvcCullData, postPass) == 132, "CullData::postPass offset");
static_assert(sizeof(NvcSceneCacheHeader) == 160, "SceneHeader must match src/scenecache.cpp:16-55");
static_assert(offsetof(NvcSceneCacheHeader, compressedVertexBytes) == 28 && offsetof(NvcSceneCacheHeader, vertexCount) == 44, "SceneHeader offsets");
static_assert(offsetof(NvcSceneCacheHeader, camera) == 108 && offsetof(NvcSceneCacheHeader, sunDirection) == 144, "SceneHeader offsets");
static_assert(sizeof(NvcKeyframe) == 32 && sizeof(NvcAnimation) == 24, "Keyframe / Animation must match src/scene.h:141-161");
#endif

#endif /* NIAGARA_CULL_H */
