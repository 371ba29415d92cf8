// nvc_api.cu — the C ABI declared in include/niagara_cull.h (context, argument validation, launches).
#include "nvc_internal.h"
#include "nvc_filter.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace
{

// Coarse Hi-Z mips staged per CTA (TMA): static + dynamic shared memory of clustercull_kernel<true, true> must stay under the
// 48 KB a kernel gets without opting in; its static part is 8.2 KB (s_stage[8][256] + barriers), so 39 KB = 9984 texels.
constexpr uint32_t kMaxStageTexels = 9984;

int cuda_fail(NvcContext* ctx, cudaError_t e, const char* what)
{
	if (ctx)
		ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
	return NVC_ERROR_CUDA;
}

// Picks the coarse tail of the pyramid that is staged into shared memory by TMA: the largest set of top mips whose
// texels fit the budget, whose start is 16-byte aligned in the packed allocation and that holds at least one 16-byte unit.
void choose_stage(nvc::HiZDesc& hz, uint32_t total_texels, uint32_t budget_texels)
{
	hz.stage_level = hz.levels;
	hz.stage_texels = 0;
	if (budget_texels == 0)
		return;
	for (uint32_t l = 0; l < hz.levels; ++l)
	{
		uint32_t tail = total_texels - hz.level_offset[l];
		uintptr_t addr = reinterpret_cast<uintptr_t>(hz.texels + hz.level_offset[l]);
		if (tail <= budget_texels && tail >= 4 && (addr & 15u) == 0)
		{
			hz.stage_level = l;
			hz.stage_texels = tail;
			return;
		}
	}
}

} // namespace

namespace nvc
{
void choose_stage_public(HiZDesc& hz, uint32_t total_texels, uint32_t budget_texels)
{
	choose_stage(hz, total_texels, budget_texels);
}
} // namespace nvc

namespace
{

// Lifetime calls (destroy / prepare / staging) work on the context's device and put the caller's device back afterwards.
struct DeviceGuard
{
	int previous = -1;
	explicit DeviceGuard(int device)
	{
		if (cudaGetDevice(&previous) != cudaSuccess)
			previous = -1;
		if (previous != device)
			cudaSetDevice(device);
		else
			previous = -1; // nothing to restore
	}
	~DeviceGuard()
	{
		if (previous >= 0)
			cudaSetDevice(previous);
	}
};

// Pass calls launch on the caller's stream, which lives on the context's device: that device must be current
// (the reference drives one VkDevice from one thread; nothing here switches devices behind the caller's back).
bool device_is_current(NvcContext* ctx)
{
	int current = -1;
	if (cudaGetDevice(&current) != cudaSuccess || current != ctx->device)
	{
		ctx->last_error = "the context's CUDA device is not the calling thread's current device";
		return false;
	}
	return true;
}

bool fill_hiz(const NvcHiZ* in, nvc::HiZDesc& out)
{
	memset(&out, 0, sizeof(out));
	if (!in)
		return false;
	if (!in->texels || in->levels == 0 || in->levels > NVC_MAX_HIZ_LEVELS || in->width == 0 || in->height == 0)
		return false;
	out.texels = in->texels;
	out.width = in->width;
	out.height = in->height;
	out.levels = in->levels;
	memcpy(out.level_offset, in->level_offset, sizeof(out.level_offset));
	out.stage_level = in->levels; // staging is opted into per pass
	out.stage_texels = 0;
	return true;
}

// Attaches the context's footprint image when it was built (nvc_depth_pyramid) for exactly this pyramid.
void attach_footprints(const NvcContext* ctx, const NvcHiZ* in, nvc::HiZDesc& out)
{
	out.fp = nullptr;
	if (!ctx->hiz_fp || !ctx->hiz_fp_valid || !in || ctx->hiz_fp_key != in->texels || ctx->hiz_fp_width != in->width || ctx->hiz_fp_height != in->height || ctx->hiz_fp_levels != in->levels)
		return;
	out.fp = ctx->hiz_fp;
	out.fp_first = ctx->hiz_fp_first;
	memcpy(out.fp_offset, ctx->hiz_fp_offset, sizeof(out.fp_offset));
}

} // namespace

extern "C"
{

NVC_API const char* nvc_version(void)
{
	return "niagara_b200 0.1 (sm_100a)";
}

NVC_API const char* nvc_status_string(int status)
{
	switch (status)
	{
	case NVC_OK:
		return "ok";
	case NVC_ERROR_INVALID_ARGUMENT:
		return "invalid argument";
	case NVC_ERROR_CUDA:
		return "CUDA error";
	case NVC_ERROR_NO_DEVICE:
		return "no CUDA device";
	case NVC_ERROR_OUT_OF_MEMORY:
		return "out of memory";
	case NVC_ERROR_NCCL:
		return "NCCL error";
	case NVC_ERROR_UNSUPPORTED:
		return "unsupported input";
	case NVC_ERROR_CORRUPT:
		return "corrupt scene cache";
	default:
		return "unknown status";
	}
}

NVC_API const char* nvc_last_error(const NvcContext* ctx)
{
	return ctx ? ctx->last_error.c_str() : "";
}

NVC_API int nvc_create(int device, const NvcLimits* limits, NvcContext** out_ctx)
{
	if (!out_ctx)
		return NVC_ERROR_INVALID_ARGUMENT;
	*out_ctx = nullptr;

	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
		return NVC_ERROR_NO_DEVICE; // no CPU fallback: the product path needs a GPU
	if (device < 0 || device >= count)
		return NVC_ERROR_INVALID_ARGUMENT;

	NvcContext* ctx = new NvcContext();
	ctx->device = device;
	if (limits)
	{
		if (limits->task_wglimit == 0 || limits->cluster_limit == 0 || limits->cluster_limit > (1u << 24) || limits->task_wglimit > (1u << 24))
		{
			delete ctx;
			return NVC_ERROR_INVALID_ARGUMENT; // cluster index packs a 24-bit command id
		}
		ctx->limits = *limits;
	}

	cudaError_t e = cudaSetDevice(device);
	if (e == cudaSuccess)
		e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
	if (e == cudaSuccess)
		e = cudaMalloc(&ctx->scratch, sizeof(nvc::Scratch));
	if (e == cudaSuccess)
		e = cudaMemset(ctx->scratch, 0, sizeof(nvc::Scratch));
	if (e == cudaSuccess)
	{
		if (const char* env = getenv("NVC_HIZ_STAGE_TEXELS"))
			ctx->hiz_stage_budget = uint32_t(strtoul(env, nullptr, 10));
		if (ctx->hiz_stage_budget > kMaxStageTexels)
			ctx->hiz_stage_budget = kMaxStageTexels;
		e = nvc::clustercull_occupancy(&ctx->cluster_blocks_early, &ctx->cluster_blocks_late, &ctx->cluster_blocks_late_staged, ctx->hiz_stage_budget * 4u);
	}
	if (e == cudaSuccess)
		e = nvc::clustercull_filter_occupancy(&ctx->cluster_filter_blocks_early, &ctx->cluster_filter_blocks_late);
	if (const char* env = getenv("NVC_CLUSTER_FILTER"))
		ctx->cluster_filter = atoi(env) != 0;
	if (const char* env = getenv("NVC_DRAW_FILTER"))
		ctx->draw_filter = atoi(env) != 0;
	if (e == cudaSuccess)
		e = cudaDeviceSynchronize();
	if (e != cudaSuccess)
	{
		fprintf(stderr, "nvc_create: %s\n", cudaGetErrorString(e));
		if (ctx->scratch)
			cudaFree(ctx->scratch);
		delete ctx;
		return e == cudaErrorMemoryAllocation ? NVC_ERROR_OUT_OF_MEMORY : NVC_ERROR_CUDA;
	}
	if (ctx->cluster_blocks_early < 1)
		ctx->cluster_blocks_early = 1;
	if (ctx->cluster_blocks_late < 1)
		ctx->cluster_blocks_late = 1;
	if (ctx->cluster_blocks_late_staged < 1)
		ctx->cluster_blocks_late_staged = 1;
	if (ctx->cluster_filter_blocks_early < 1)
		ctx->cluster_filter_blocks_early = 1;
	if (ctx->cluster_filter_blocks_late < 1)
		ctx->cluster_filter_blocks_late = 1;

	*out_ctx = ctx;
	return NVC_OK;
}

NVC_API void nvc_destroy(NvcContext* ctx)
{
	if (!ctx)
		return;
	DeviceGuard guard(ctx->device);
	nvc::gather_destroy(ctx);
	nvc::nccl_destroy(ctx);
	if (ctx->scratch)
		cudaFree(ctx->scratch);
	if (ctx->mesh_heads)
		cudaFree(ctx->mesh_heads);
	if (ctx->mesh_errors)
		cudaFree(ctx->mesh_errors);
	if (ctx->hiz_fp)
		cudaFree(ctx->hiz_fp);
	delete ctx;
}

NVC_API int nvc_prepare_meshes(NvcContext* ctx, void* stream, const NvcMesh* meshes, uint32_t mesh_count)
{
	if (!ctx)
		return NVC_ERROR_INVALID_ARGUMENT;
	DeviceGuard guard(ctx->device);
	if (ctx->mesh_heads)
	{
		cudaDeviceSynchronize();
		cudaFree(ctx->mesh_heads);
		cudaFree(ctx->mesh_errors);
		ctx->mesh_heads = nullptr;
		ctx->mesh_errors = nullptr;
		ctx->prepared_meshes = nullptr;
		ctx->prepared_mesh_count = 0;
	}
	if (!meshes || mesh_count == 0)
		return NVC_OK; // un-prepare
	cudaError_t e = cudaMalloc(&ctx->mesh_heads, size_t(mesh_count) * sizeof(nvc::MeshCullHead));
	if (e == cudaSuccess)
		e = cudaMalloc(&ctx->mesh_errors, size_t(mesh_count) * NVC_MAX_LODS * sizeof(float));
	if (e == cudaSuccess)
		e = nvc::launch_pack_meshes(meshes, mesh_count, static_cast<nvc::MeshCullHead*>(ctx->mesh_heads), ctx->mesh_errors, static_cast<cudaStream_t>(stream));
	if (e != cudaSuccess)
	{
		cudaFree(ctx->mesh_heads);
		cudaFree(ctx->mesh_errors);
		ctx->mesh_heads = nullptr;
		ctx->mesh_errors = nullptr;
		return e == cudaErrorMemoryAllocation ? NVC_ERROR_OUT_OF_MEMORY : cuda_fail(ctx, e, "nvc_prepare_meshes");
	}
	ctx->prepared_meshes = meshes;
	ctx->prepared_mesh_count = mesh_count;
	return NVC_OK;
}

NVC_API int nvc_set_hiz_staging(NvcContext* ctx, uint32_t texels)
{
	if (!ctx)
		return NVC_ERROR_INVALID_ARGUMENT;
	DeviceGuard guard(ctx->device);
	ctx->hiz_stage_budget = texels > kMaxStageTexels ? kMaxStageTexels : texels;
	int early = 0, late = 0;
	cudaError_t e = nvc::clustercull_occupancy(&early, &late, &ctx->cluster_blocks_late_staged, ctx->hiz_stage_budget * 4u);
	if (e != cudaSuccess)
		return cuda_fail(ctx, e, "nvc_set_hiz_staging");
	if (ctx->cluster_blocks_late_staged < 1)
		ctx->cluster_blocks_late_staged = 1;
	return NVC_OK;
}

NVC_API int nvc_set_cluster_filter(NvcContext* ctx, int enabled)
{
	if (!ctx)
		return NVC_ERROR_INVALID_ARGUMENT;
	ctx->cluster_filter = enabled != 0;
	return NVC_OK;
}

NVC_API int nvc_filter_stats(NvcContext* ctx, uint64_t* out_items_undecided2, int reset)
{
	if (!ctx)
		return NVC_ERROR_INVALID_ARGUMENT;
	DeviceGuard guard(ctx->device);
	cudaError_t e = cudaDeviceSynchronize();
	unsigned long long v[2] = { 0, 0 };
	if (e == cudaSuccess)
		e = cudaMemcpy(v, &ctx->scratch->filter_items, sizeof(v), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && reset)
		e = cudaMemset(&ctx->scratch->filter_items, 0, sizeof(v));
	if (e != cudaSuccess)
		return cuda_fail(ctx, e, "nvc_filter_stats");
	if (out_items_undecided2)
	{
		out_items_undecided2[0] = v[0];
		out_items_undecided2[1] = v[1];
	}
	return NVC_OK;
}

NVC_API int nvc_drawcull(NvcContext* ctx, void* stream, const NvcCullData* cull, int late, int task,
    const NvcMeshDraw* draws, const NvcMesh* meshes, uint32_t* draw_visibility,
    void* commands, uint32_t* command_count4, const NvcHiZ* hiz)
{
	if (!ctx || !cull || !commands || !command_count4)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (cull->drawCount && (!draws || !meshes || !draw_visibility))
		return NVC_ERROR_INVALID_ARGUMENT;
	if (!device_is_current(ctx))
		return NVC_ERROR_INVALID_ARGUMENT;

	nvc::DrawCullParams p;
	memset(&p, 0, sizeof(p));
	p.cull = *cull;
	p.draws = draws;
	p.meshes = meshes;
	if (ctx->mesh_heads && ctx->prepared_meshes == meshes)
	{
		p.mesh_heads = static_cast<const nvc::MeshCullHead*>(ctx->mesh_heads);
		p.mesh_errors = ctx->mesh_errors;
	}
	p.draw_visibility = draw_visibility;
	p.commands = commands;
	p.command_count4 = command_count4;
	p.scratch = ctx->scratch;
	p.task_wglimit = ctx->limits.task_wglimit;
	if (late && task && ctx->gather)
		p.mc_commands = nvc::gather_fused_target(ctx); // armed by nvc_gather_fuse_next_drawcull, else nullptr
	bool need_hiz = late && cull->occlusionEnabled == 1;
	if (!fill_hiz(hiz, p.hiz) && need_hiz)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (need_hiz && ctx->cluster_filter && ctx->draw_filter)
	{
		// the occlusion stage as a conservative filter (same switch and the same validity domain as the cluster pass)
		attach_footprints(ctx, hiz, p.hiz);
		p.filter = nvc::make_filter_consts(p.cull, p.hiz, true);
		p.use_filter = (p.filter.enabled && p.filter.occ_ok) ? 1u : 0u;
	}

	cudaError_t e = nvc::launch_drawcull(p, late != 0, task != 0, static_cast<cudaStream_t>(stream));
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_drawcull");
}

static int fill_cluster_params(NvcContext* ctx, const NvcCullData* cull, int late, const NvcMeshTaskCommand* task_commands,
    const uint32_t* command_count4, const NvcMeshDraw* draws, const NvcMeshlet* meshlets, uint32_t* meshlet_visibility,
    const NvcHiZ* hiz, nvc::ClusterParams& p)
{
	if (!ctx || !cull || !task_commands || !command_count4 || !draws || !meshlets)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (cull->clusterOcclusionEnabled == 1 && !meshlet_visibility)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (!device_is_current(ctx))
		return NVC_ERROR_INVALID_ARGUMENT;
	memset(&p, 0, sizeof(p));
	p.cull = *cull;
	p.task_commands = task_commands;
	p.command_count4 = command_count4;
	p.draws = draws;
	p.meshlets = meshlets;
	p.meshlet_visibility = meshlet_visibility;
	p.scratch = ctx->scratch;
	p.cluster_limit = ctx->limits.cluster_limit;
	p.one = 1.0f;
	p.neg_one = -1.0f;
	bool need_hiz = late && cull->clusterOcclusionEnabled == 1;
	if (!fill_hiz(hiz, p.hiz) && need_hiz)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (need_hiz)
		attach_footprints(ctx, hiz, p.hiz);
	return NVC_OK;
}

NVC_API int nvc_clustercull(NvcContext* ctx, void* stream, const NvcCullData* cull, int late,
    const NvcMeshTaskCommand* task_commands, const uint32_t* command_count4,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, uint32_t* meshlet_visibility,
    uint32_t* cluster_indices, uint32_t* cluster_count4, const NvcHiZ* hiz)
{
	if (!cluster_indices || !cluster_count4)
		return NVC_ERROR_INVALID_ARGUMENT;
	nvc::ClusterParams p;
	int s = fill_cluster_params(ctx, cull, late, task_commands, command_count4, draws, meshlets, meshlet_visibility, hiz, p);
	if (s != NVC_OK)
		return s;
	p.cluster_indices = cluster_indices;
	p.cluster_count4 = cluster_count4;

	bool staged = false;
	if (late && cull->clusterOcclusionEnabled == 1 && hiz)
	{
		nvc::choose_stage_public(p.hiz, hiz->total_texels, ctx->hiz_stage_budget);
		staged = p.hiz.stage_level < p.hiz.levels;
	}
	uint32_t blocks = uint32_t(ctx->sm_count) * uint32_t(late ? (staged ? ctx->cluster_blocks_late_staged : ctx->cluster_blocks_late) : ctx->cluster_blocks_early);
	// default: the filtered kernel (conservative filter + exact fallback, same results); TMA-staged Hi-Z and
	// nvc_set_cluster_filter(ctx, 0) select the exact kernel
	if (ctx->cluster_filter && !staged)
	{
		const bool need_occlusion = late && cull->clusterOcclusionEnabled == 1;
		p.filter = nvc::make_filter_consts(p.cull, p.hiz, need_occlusion);
		// unusual view / projection / pyramid shapes are outside the filter's error analysis: the exact kernel takes the pass
		if (p.filter.enabled && (!need_occlusion || p.filter.occ_ok))
		{
			p.use_filter = 1;
			blocks = uint32_t(ctx->sm_count) * uint32_t(late ? ctx->cluster_filter_blocks_late : ctx->cluster_filter_blocks_early);
		}
	}
	// multi-GPU: the persistent grid would otherwise own every register of every SM for the whole pass, and the exchange's
	// one-block helper kernels (acknowledgement wait, flag raise) could not start before it ends — leave two CTA slots free
	if (ctx->gather)
	{
		const uint32_t reserve = nvc::gather_reserved_blocks(ctx);
		if (blocks > 4 * reserve)
			blocks -= reserve;
	}
	cudaError_t e = nvc::launch_clustercull(p, late != 0, blocks, static_cast<cudaStream_t>(stream));
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_clustercull");
}

NVC_API int nvc_taskcull(NvcContext* ctx, void* stream, const NvcCullData* cull, int late,
    const NvcMeshTaskCommand* task_commands, const uint32_t* command_count4,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, uint32_t* meshlet_visibility,
    NvcMeshTaskPayload* payloads, uint32_t* emit_counts, const NvcHiZ* hiz)
{
	if (!payloads || !emit_counts)
		return NVC_ERROR_INVALID_ARGUMENT;
	nvc::ClusterParams p;
	int s = fill_cluster_params(ctx, cull, late, task_commands, command_count4, draws, meshlets, meshlet_visibility, hiz, p);
	if (s != NVC_OK)
		return s;
	uint32_t blocks = uint32_t(ctx->sm_count) * 4u;
	if (ctx->cluster_filter)
	{
		const bool need_occlusion = late && cull->clusterOcclusionEnabled == 1;
		p.filter = nvc::make_filter_consts(p.cull, p.hiz, need_occlusion);
		if (p.filter.enabled && (!need_occlusion || p.filter.occ_ok))
		{
			p.use_filter = 1;
			blocks = uint32_t(ctx->sm_count) * uint32_t(late ? ctx->cluster_filter_blocks_late : ctx->cluster_filter_blocks_early);
		}
	}
	cudaError_t e = nvc::launch_taskcull(p, late != 0, payloads, emit_counts, blocks, static_cast<cudaStream_t>(stream));
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_taskcull");
}

NVC_API int nvc_decode_clusters(NvcContext* ctx, void* stream, const uint32_t* cluster_indices, const uint32_t* cluster_count4,
    const NvcMeshTaskCommand* task_commands, const NvcMeshlet* meshlets, NvcClusterRecord* records, uint32_t* stats4)
{
	if (!ctx || !cluster_indices || !cluster_count4 || !task_commands || !meshlets || !stats4)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (!device_is_current(ctx))
		return NVC_ERROR_INVALID_ARGUMENT;
	cudaError_t e = nvc::launch_decode_clusters(cluster_indices, cluster_count4, task_commands, meshlets, records, stats4, uint32_t(ctx->sm_count) * 8u, static_cast<cudaStream_t>(stream));
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_decode_clusters");
}

NVC_API int nvc_raster_depth(NvcContext* ctx, void* stream, const float* projection16, const NvcCullData* pass, const uint32_t* cluster_indices, const uint32_t* cluster_count4,
    const NvcMeshTaskCommand* task_commands, const NvcMeshDraw* draws, const NvcMeshlet* meshlets, const uint32_t* meshletdata, uint32_t meshletdata_words, const NvcVertex* vertices,
    uint32_t vertex_count, float* depth, uint32_t width, uint32_t height, uint32_t* stats4)
{
	if (!ctx || !projection16 || !pass || !cluster_indices || !cluster_count4 || !task_commands || !draws || !meshlets || !meshletdata || !vertices || !depth || width == 0 || height == 0)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (uint64_t(width) * height >= (1ull << 31))
		return NVC_ERROR_UNSUPPORTED;
	if (!device_is_current(ctx))
		return NVC_ERROR_INVALID_ARGUMENT;
	cudaError_t e = nvc::launch_raster_depth(projection16, *pass, cluster_indices, cluster_count4, task_commands, draws, meshlets, meshletdata, meshletdata_words, vertices, vertex_count, depth,
	    width, height, stats4, uint32_t(ctx->sm_count) * 16u, static_cast<cudaStream_t>(stream));
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_raster_depth");
}

NVC_API int nvc_depth_pyramid(NvcContext* ctx, void* stream, const float* depth,
    uint32_t depth_width, uint32_t depth_height, const NvcHiZ* hiz)
{
	if (!ctx || !depth || !hiz || depth_width == 0 || depth_height == 0)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (!device_is_current(ctx))
		return NVC_ERROR_INVALID_ARGUMENT;
	nvc::PyramidParams p;
	memset(&p, 0, sizeof(p));
	if (!fill_hiz(hiz, p.hiz))
		return NVC_ERROR_INVALID_ARGUMENT;
	// the pyramid must be the one the reference would allocate for this depth target (niagara.cpp:1339-1342)
	if (hiz->width != nvc_previous_pow2(depth_width) || hiz->height != nvc_previous_pow2(depth_height) ||
	    hiz->levels != nvc_image_mip_levels(hiz->width, hiz->height))
		return NVC_ERROR_INVALID_ARGUMENT;
	p.depth = depth;
	p.depth_width = depth_width;
	p.depth_height = depth_height;
	p.scratch = ctx->scratch;
	cudaError_t e = nvc::launch_pyramid(p, static_cast<cudaStream_t>(stream));
	if (e == cudaSuccess && ctx->hiz_fp && ctx->hiz_fp_key == hiz->texels && ctx->hiz_fp_width == hiz->width && ctx->hiz_fp_height == hiz->height && ctx->hiz_fp_levels == hiz->levels)
	{
		// derived footprint image of the pyramid just built (nvc_prepare_hiz), one more launch on the same stream
		memcpy(p.hiz.fp_offset, ctx->hiz_fp_offset, sizeof(p.hiz.fp_offset));
		p.hiz.fp_first = ctx->hiz_fp_first;
		e = nvc::launch_footprint(p.hiz, ctx->hiz_fp, ctx->hiz_fp_total, static_cast<cudaStream_t>(stream));
		ctx->hiz_fp_valid = e == cudaSuccess;
	}
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_depth_pyramid");
}

NVC_API int nvc_prepare_hiz(NvcContext* ctx, const NvcHiZ* hiz)
{
	if (!ctx)
		return NVC_ERROR_INVALID_ARGUMENT;
	DeviceGuard guard(ctx->device);
	if (ctx->hiz_fp)
	{
		cudaDeviceSynchronize();
		cudaFree(ctx->hiz_fp);
		ctx->hiz_fp = nullptr;
	}
	ctx->hiz_fp_key = nullptr;
	ctx->hiz_fp_valid = false;
	ctx->hiz_fp_total = 0;
	if (!hiz)
		return NVC_OK; // un-prepare
	if (!hiz->texels || hiz->levels == 0 || hiz->levels > NVC_MAX_HIZ_LEVELS || hiz->width == 0 || hiz->height == 0 || hiz->width > 65536 || hiz->height > 65536)
		return NVC_ERROR_INVALID_ARGUMENT;
	uint64_t total = 0;
	if (const char* env = getenv("NVC_FP_FIRST_LEVEL"))
		ctx->hiz_fp_first = uint32_t(strtoul(env, nullptr, 10));
	if (ctx->hiz_fp_first >= hiz->levels)
		ctx->hiz_fp_first = hiz->levels - 1; // at least the coarsest mip has an image
	for (uint32_t l = 0; l < hiz->levels; ++l)
	{
		uint32_t w = hiz->width >> l, h = hiz->height >> l;
		w = w ? w : 1;
		h = h ? h : 1;
		ctx->hiz_fp_offset[l] = uint32_t(total);
		if (l >= ctx->hiz_fp_first)
			total += uint64_t(nvc::fp_pitch(w)) * (h + 1);
	}
	if (total >= (1ull << 31))
		return NVC_ERROR_UNSUPPORTED;
	cudaError_t e = cudaMalloc(&ctx->hiz_fp, size_t(total) * sizeof(float));
	if (e != cudaSuccess)
	{
		ctx->hiz_fp = nullptr;
		return e == cudaErrorMemoryAllocation ? NVC_ERROR_OUT_OF_MEMORY : cuda_fail(ctx, e, "nvc_prepare_hiz");
	}
	ctx->hiz_fp_key = hiz->texels;
	ctx->hiz_fp_width = hiz->width;
	ctx->hiz_fp_height = hiz->height;
	ctx->hiz_fp_levels = hiz->levels;
	ctx->hiz_fp_total = uint32_t(total);
	return NVC_OK;
}

NVC_API int nvc_hiz_footprints(NvcContext* ctx, const float** image_out, uint32_t* first_level_out, uint32_t* offset_out, uint32_t* total_out)
{
	if (!ctx || !image_out)
		return NVC_ERROR_INVALID_ARGUMENT;
	const bool have = ctx->hiz_fp && ctx->hiz_fp_valid;
	*image_out = have ? ctx->hiz_fp : nullptr;
	if (first_level_out)
		*first_level_out = ctx->hiz_fp_first;
	if (offset_out)
		memcpy(offset_out, ctx->hiz_fp_offset, sizeof(ctx->hiz_fp_offset));
	if (total_out)
		*total_out = have ? ctx->hiz_fp_total : 0u;
	return NVC_OK;
}

NVC_API int nvc_update_draws(NvcContext* ctx, void* stream, NvcMeshDraw* draws, uint32_t draw_count, const uint32_t* update_indices,
    const NvcMeshDraw* update_values, uint32_t count)
{
	if (!ctx || !draws || (count && (!update_indices || !update_values)))
		return NVC_ERROR_INVALID_ARGUMENT;
	if ((reinterpret_cast<uintptr_t>(draws) | reinterpret_cast<uintptr_t>(update_values)) & 15u)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (!device_is_current(ctx))
		return NVC_ERROR_INVALID_ARGUMENT;
	cudaError_t e = nvc::launch_update_draws(draws, draw_count, update_indices, update_values, count, static_cast<cudaStream_t>(stream));
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_update_draws");
}

NVC_API int nvc_cook_meshlet_bounds(NvcContext* ctx, void* stream, const NvcVertex* vertices, uint32_t vertex_count, const uint32_t* meshletdata,
    uint32_t meshletdata_words, NvcMeshlet* meshlets, uint32_t meshlet_count, uint32_t* rejected)
{
	if (!ctx || (meshlet_count && (!vertices || !meshletdata || !meshlets)))
		return NVC_ERROR_INVALID_ARGUMENT;
	if (!device_is_current(ctx))
		return NVC_ERROR_INVALID_ARGUMENT;
	cudaError_t e = nvc::launch_cook_meshlet_bounds(vertices, vertex_count, meshletdata, meshletdata_words, meshlets, meshlet_count, rejected, static_cast<cudaStream_t>(stream));
	return e == cudaSuccess ? NVC_OK : cuda_fail(ctx, e, "nvc_cook_meshlet_bounds");
}

} // extern "C"
