// TEST INFRASTRUCTURE — scheduler of the CPU SIMT emulation (see include/cuda_runtime.h in this directory).
#include "cuda_runtime.h"

#include <stdio.h>
#include <ucontext.h>

#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace
{

struct Fiber
{
	ucontext_t ctx;
	std::vector<char> stack;
	bool done = false;
	uint3 tid;
	uint32_t linear = 0;
};

struct Warp
{
	uint32_t values[32];
	uint32_t snapshot[32];
	uint32_t arrived = 0; // lanes waiting in the current rendezvous
	uint32_t pending = 0; // lanes that still have to pick up the snapshot
	uint32_t expected = 0;
	bool releasing = false;
};

ucontext_t g_scheduler;
std::vector<Fiber> g_fibers;
std::vector<Warp> g_warps;
Fiber* g_current = nullptr;
const std::function<void()>* g_body = nullptr;
std::vector<char> g_dynamic;
uint32_t g_live = 0, g_bar_count = 0, g_bar_generation = 0;
uint64_t g_progress = 0; // bumped whenever any fiber gets past a rendezvous or finishes: deadlock detection

void yield()
{
	swapcontext(&g_current->ctx, &g_scheduler);
}

void fiberEntry()
{
	(*g_body)();
	g_current->done = true;
	--g_live;
	++g_progress;
	if (g_bar_count && g_bar_count == g_live) // the others are waiting in __syncthreads for a thread that has now exited
	{
		g_bar_count = 0;
		++g_bar_generation;
	}
	swapcontext(&g_current->ctx, &g_scheduler);
}

} // namespace

namespace emu
{

uint32_t lane()
{
	return g_current->linear & 31u;
}

void* dynamic_smem()
{
	return g_dynamic.data();
}

const uint32_t* warp_gather(uint32_t mask, uint32_t value)
{
	Warp& w = g_warps[g_current->linear >> 5];
	const uint32_t bit = 1u << lane();
	while (w.releasing) // the previous rendezvous is still handing out its snapshot
		yield();
	w.values[lane()] = value;
	w.arrived |= bit;
	w.expected |= mask;
	if ((w.arrived & w.expected) == w.expected)
	{
		memcpy(w.snapshot, w.values, sizeof(w.values));
		w.releasing = true;
		w.pending = w.arrived;
	}
	else
		while (!(w.releasing && (w.pending & bit)))
			yield();
	w.pending &= ~bit;
	++g_progress;
	if (w.pending == 0)
	{
		w.arrived = 0;
		w.expected = 0;
		w.releasing = false;
	}
	return w.snapshot; // valid until this fiber yields again (single OS thread)
}

void block_barrier()
{
	uint32_t generation = g_bar_generation;
	if (++g_bar_count == g_live)
	{
		g_bar_count = 0;
		++g_bar_generation;
	}
	else
		while (g_bar_generation == generation)
			yield();
	++g_progress;
}

void launch(dim3 grid, dim3 block, size_t dynamic_bytes, const std::function<void()>& body)
{
	const uint32_t threads = block.x * block.y * block.z;
	gridDim = grid;
	blockDim = block;
	g_body = &body;
	g_dynamic.assign(dynamic_bytes + 16, 0);
	g_fibers.resize(threads);
	for (uint32_t bz = 0; bz < grid.z; ++bz)
		for (uint32_t by = 0; by < grid.y; ++by)
			for (uint32_t bx = 0; bx < grid.x; ++bx)
			{
				g_warps.assign((threads + 31) / 32, Warp());
				g_live = threads;
				g_bar_count = 0;
				for (uint32_t t = 0; t < threads; ++t)
				{
					Fiber& f = g_fibers[t];
					f.done = false;
					f.linear = t;
					f.tid = uint3{ t % block.x, (t / block.x) % block.y, t / (block.x * block.y) };
					if (f.stack.empty())
						f.stack.resize(192 * 1024);
					getcontext(&f.ctx);
					f.ctx.uc_stack.ss_sp = f.stack.data();
					f.ctx.uc_stack.ss_size = f.stack.size();
					f.ctx.uc_link = nullptr;
					makecontext(&f.ctx, fiberEntry, 0);
				}
				while (g_live)
				{
					uint64_t before = g_progress;
					for (uint32_t t = 0; t < threads; ++t)
					{
						Fiber& f = g_fibers[t];
						if (f.done)
							continue;
						g_current = &f;
						threadIdx = f.tid;
						blockIdx = uint3{ bx, by, bz };
						swapcontext(&g_scheduler, &f.ctx);
					}
					if (g_live && g_progress == before)
					{
						fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %u threads wait for a lane / thread that never arrives\n", bx, by, bz, g_live);
						abort();
					}
				}
			}
	g_body = nullptr;
}

} // namespace emu
