// nvc_internal.h — types shared between the C ABI (nvc_api.cu) and the kernels (nvc_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/niagara_cull.h"

namespace nvc
{
struct Scratch;
}

struct NvcContext
{
	int device = 0;
	int sm_count = 0;
	int cluster_blocks_early = 0, cluster_blocks_late = 0, cluster_blocks_late_staged = 0;
	int cluster_filter_blocks_early = 0, cluster_filter_blocks_late = 0;
	bool draw_filter = false;   // env NVC_DRAW_FILTER=1: filtered occlusion stage of the late drawcull (measured: no gain on B200, the pass is not issue bound)
	bool cluster_filter = true; // nvc_set_cluster_filter / env NVC_CLUSTER_FILTER: filtered cluster kernel (default) or the exact one
	uint32_t hiz_stage_budget = 0; // texels (24 KB) of coarse Hi-Z mips staged per CTA; 0 = off (env NVC_HIZ_STAGE_TEXELS)
	NvcLimits limits = { NVC_TASK_WGLIMIT, NVC_CLUSTER_LIMIT };
	nvc::Scratch* scratch = nullptr;
	std::string last_error;
	void* nccl_comm = nullptr; // ncclComm_t, owned (nvc_nccl.cpp)
	void* gather = nullptr;    // NvcGather, owned (nvc_peer.cu)
	// derived mesh cull view (nvc_prepare_meshes)
	const void* prepared_meshes = nullptr;
	uint32_t prepared_mesh_count = 0;
	void* mesh_heads = nullptr;
	float* mesh_errors = nullptr;
	int nccl_rank = 0, nccl_world = 1;
	// footprint image of one pyramid (nvc_prepare_hiz): keyed by the pyramid's texel pointer and shape
	float* hiz_fp = nullptr;
	const float* hiz_fp_key = nullptr;
	uint32_t hiz_fp_width = 0, hiz_fp_height = 0, hiz_fp_levels = 0, hiz_fp_total = 0;
	uint32_t hiz_fp_first = 0; // first mip with a footprint image (env NVC_FP_FIRST_LEVEL)
	uint32_t hiz_fp_offset[NVC_MAX_HIZ_LEVELS] = {};
	bool hiz_fp_valid = false; // set by nvc_depth_pyramid, cleared by nvc_prepare_hiz
};

namespace nvc
{
void nccl_destroy(NvcContext* ctx);   // nvc_nccl.cpp
void gather_destroy(NvcContext* ctx); // nvc_peer.cu
uint32_t* gather_fused_target(NvcContext* ctx);
uint32_t gather_reserved_blocks(NvcContext* ctx); // nvc_peer.cu: CTA slots the persistent cluster grid leaves to the exchange's kernels // nvc_peer.cu: multicast destination of the armed late drawcull, or nullptr (disarms)

// Device-side counters owned by the context.  Each pass's last-block epilogue leaves them zeroed, which replaces
// the reference's vkCmdFillBuffer resets (niagara.cpp:1541,1586) and keeps every pass a single launch.
// Hot atomics live on separate 128-byte lines.
struct alignas(128) Scratch
{
	uint32_t draw_counter;
	uint32_t pad0[31];
	uint32_t draw_done;
	uint32_t pad1[31];
	uint32_t cluster_counter;
	uint32_t pad2[31];
	uint32_t cluster_done;
	uint32_t pad3[31];
	uint32_t cluster_batch;
	uint32_t pad4[31];
	uint32_t pyramid_done;
	uint32_t pad5[31];
	// filtered cluster pass, cumulative since the last nvc_filter_stats(reset): meshlets evaluated, meshlets that took the exact path
	unsigned long long filter_items, filter_undecided;
	uint32_t pad6[28];
};

// row pitch of a footprint-image level of width w (w + 1 entries per row, padded so that every row starts 16-byte aligned)
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t fp_pitch(uint32_t w)
{
	return (w + 4u) & ~3u;
}

struct HiZDesc
{
	float* texels;
	uint32_t width, height, levels;
	uint32_t level_offset[NVC_MAX_HIZ_LEVELS];
	// mips >= stage_level (the coarse tail of the packed pyramid, stage_texels texels) are staged into shared memory
	// once per CTA with one TMA bulk copy; stage_level == levels disables staging
	uint32_t stage_level, stage_texels;
	// Footprint image (nvc_prepare_hiz; built by nvc_depth_pyramid): for every mip, F(i, j) = min of the 2 x 2 texel
	// footprint whose lower corner is (i, j), i in [-1, w-1], j in [-1, h-1], indices clamped to the level — what the
	// MIN-reduction sampler returns for any coordinate whose bilinear footprint starts there (resources.cpp:294-325).
	// Level l: fp + fp_offset[l], pitch w + 1, entry (i + 1, j + 1).  fp == nullptr: not available.
	// Level l: fp + fp_offset[l], row pitch fp_pitch(w) (a multiple of four entries: rows are 16-byte aligned), entry (i + 1, j + 1).
	// Only mips >= fp_first have an image (the finest mips hold 15/16 of the texels and are sampled by sub-8-pixel spheres only;
	// lookups into them take the four texel loads).
	const float* fp;
	uint32_t fp_offset[NVC_MAX_HIZ_LEVELS];
	uint32_t fp_first;
};

// Derived, read-only cull view of the reference's 208-byte Mesh records, built once per geometry upload by
// nvc_prepare_meshes: one 32-byte head per mesh (everything a single-LOD / LOD-0 task draw needs = ONE sector instead
// of the 3-4 sectors the AoS struct spreads it over) and the 8 LOD errors as a second 32-byte record.
struct MeshCullHead
{
	float center[3];
	float radius;
	uint32_t lodCount;
	uint32_t lod0MeshletOffset;
	uint32_t lod0MeshletCount;
	uint32_t vertexOffset;
};
static_assert(sizeof(MeshCullHead) == 32, "one sector");

// Per-launch constants of the conservative meshlet filter (nvc_filter.cuh; host-computed by make_filter_consts).
// Grouped in 16-byte vectors in the order the kernel consumes them, so that each group is ONE uniform load per chunk.
struct FilterConsts
{
	float4 fr;        // mFk (frustum margin = mFk * E), zfarLo, zfarHi (zfar (1 -+ 2^-20)), zn4u (4 u znear, enters Et)
	float4 pr;        // hPx, hPyn (0.5 P00, -0.5 P11), sxk, syk (size.x * pyramidWidth = r vx icz sxk, sxk = 2 hPx pw; same for y)
	float4 cg;        // kGx, kGrx, kGy, kGry: validity cone per axis |cx| + r kGrx <= kGx cz (kGx = 1 / hPx, kGrx = sqrt(1 + kGx^2))
	float4 mk;        // Km1, Km2 (dm = gr (Km1 + Km2 m)), KuvP (max(pw, ph) Kuv 1.05), Kfp (footprint margin = Kfp whf gr)
	uint4 lv;         // float bits of 2^levels, 2^(levels-1), (float)hiz.width, (float)hiz.height
	float vrE;        // max(max row abs sum of V3, 1)
	float Kuv;        // uv error = Kuv g relE (informational; folded into KuvP / Kfp)
	uint32_t enabled; // 0: every item is undecided (unusual view / projection: the exact path does everything)
	uint32_t occ_ok;  // 0: the occlusion stage is never decided here (non power-of-two pyramid, ...)
};

struct DrawCullParams
{
	NvcCullData cull;
	const NvcMeshDraw* draws;
	const NvcMesh* meshes;
	const MeshCullHead* mesh_heads; // may be null: read everything from `meshes`
	const float* mesh_errors;       // [mesh][8], valid when mesh_heads != null
	uint32_t* draw_visibility;
	void* commands;
	uint32_t* command_count4;
	Scratch* scratch;
	HiZDesc hiz;
	uint32_t task_wglimit;
	// multi-GPU, fused all-gather (nvc_gather_fuse_next_drawcull): NVSwitch multicast alias of THIS rank's slot in every
	// rank's gathered slab buffer; every command the pass writes to `commands` is also stored there (replicated by the
	// switch to all ranks), so the exchange happens inside the producing kernel.  nullptr: not fused.
	uint32_t* mc_commands;
	// late pass: the occlusion stage runs as a conservative filter on the exact centre; undecided draws take the exact path
	FilterConsts filter;
	uint32_t use_filter;
};

struct ClusterParams
{
	NvcCullData cull;
	const NvcMeshTaskCommand* task_commands;
	const uint32_t* command_count4;
	const NvcMeshDraw* draws;
	const NvcMeshlet* meshlets;
	uint32_t* meshlet_visibility;
	uint32_t* cluster_indices;
	uint32_t* cluster_count4;
	Scratch* scratch;
	HiZDesc hiz;
	uint32_t cluster_limit;
	// task-shading submission mode (meshlet.task.glsl): per-command payloads + emit counts instead of cib / ccb
	NvcMeshTaskPayload* payloads;
	uint32_t* emit_counts;
	FilterConsts filter; // valid when use_filter
	uint32_t use_filter; // 1: clustercull_filter_kernel (conservative filter + exact fallback), 0: the exact kernel
	float one, neg_one; // 1.0f / -1.0f as run-time values: see nvc_math2.cuh (keeps ptxas from contracting packed adds)
};

struct PyramidParams
{
	const float* depth;
	uint32_t depth_width, depth_height;
	HiZDesc hiz;
	Scratch* scratch;
	uint32_t vector_ok; // set by launch_pyramid: bases aligned for 16-byte loads
};

cudaError_t launch_drawcull(const DrawCullParams& p, bool late, bool task, cudaStream_t stream);
cudaError_t launch_clustercull(const ClusterParams& p, bool late, uint32_t blocks, cudaStream_t stream);
cudaError_t launch_taskcull(const ClusterParams& p, bool late, NvcMeshTaskPayload* payloads, uint32_t* emit_counts, uint32_t blocks, cudaStream_t stream);
cudaError_t launch_pyramid(const PyramidParams& p, cudaStream_t stream);
cudaError_t launch_footprint(const HiZDesc& hiz, float* fp, uint32_t total, cudaStream_t stream);
cudaError_t launch_decode_clusters(const uint32_t* cluster_indices, const uint32_t* cluster_count4, const NvcMeshTaskCommand* task_commands, const NvcMeshlet* meshlets, NvcClusterRecord* records, uint32_t* stats4, uint32_t blocks, cudaStream_t stream);
cudaError_t launch_raster_depth(const float* projection16, const NvcCullData& pass, const uint32_t* cluster_indices, const uint32_t* cluster_count4, const NvcMeshTaskCommand* task_commands,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, const uint32_t* meshletdata, uint32_t meshletdata_words, const NvcVertex* vertices, uint32_t vertex_count, float* depth, uint32_t width,
    uint32_t height, uint32_t* stats4, uint32_t blocks, cudaStream_t stream);
cudaError_t launch_cook_meshlet_bounds(const NvcVertex* vertices, uint32_t vertex_count, const uint32_t* meshletdata, uint32_t meshletdata_words, NvcMeshlet* meshlets, uint32_t meshlet_count, uint32_t* rejected, cudaStream_t stream);
cudaError_t launch_update_draws(NvcMeshDraw* draws, uint32_t draw_count, const uint32_t* update_indices, const NvcMeshDraw* update_values, uint32_t count, cudaStream_t stream);
cudaError_t launch_pack_meshes(const NvcMesh* meshes, uint32_t count, MeshCullHead* heads, float* errors, cudaStream_t stream);
cudaError_t clustercull_occupancy(int* blocks_per_sm_early, int* blocks_per_sm_late, int* blocks_per_sm_late_staged, uint32_t stage_bytes);
cudaError_t clustercull_filter_occupancy(int* blocks_per_sm_early, int* blocks_per_sm_late);
uint32_t hiz_stage_bytes(const HiZDesc& hiz);
void choose_stage_public(HiZDesc& hz, uint32_t total_texels, uint32_t budget_texels);

} // namespace nvc
