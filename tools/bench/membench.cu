// membench.cu — what the streams of the cull passes can reach on this GPU (read-only, DRAM-resident inputs):
//   a: every byte, 16-byte loads, one pass           (copy-kernel style upper bound for reads)
//   b: 8 + 4 bytes of every 24-byte record (the cluster passes' Meshlet access), thread per record, one wave after another
//   c: the same records by persistent warps, 32 consecutive records per step, `depth` steps of loads in flight
//   d: 3 x 16 bytes of every 48-byte record (MeshDraw), thread per record
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu ; run: ./membench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_full(const uint4* __restrict__ p, size_t n, unsigned* out)
{
	unsigned acc = 0;
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
	{
		uint4 v = __ldg(p + i);
		acc += v.x ^ v.y ^ v.z ^ v.w;
	}
	if (acc == 0x12345678u)
		*out = acc;
}

__global__ void k_rec24(const char* __restrict__ p, size_t n, unsigned* out)
{
	unsigned acc = 0;
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
	{
		const char* r = p + i * 24;
		uint2 a = __ldg(reinterpret_cast<const uint2*>(r));
		unsigned b = __ldg(reinterpret_cast<const unsigned*>(r + 8));
		acc += a.x ^ a.y ^ b;
	}
	if (acc == 0x12345678u)
		*out = acc;
}

template <int DEPTH>
__global__ void k_rec24_persist(const char* __restrict__ p, size_t n, unsigned* ticket, unsigned* out)
{
	// a warp takes 320 consecutive records (10 steps of 32) per ticket, like a batch of the cluster pass
	const unsigned lane = threadIdx.x & 31u;
	unsigned acc = 0;
	const size_t nbatch = n / 320;
	for (;;)
	{
		unsigned b = 0;
		if (lane == 0)
			b = atomicAdd(ticket, 1u);
		b = __shfl_sync(0xffffffffu, b, 0);
		if (b >= nbatch)
			break;
		const char* base = p + size_t(b) * 320 * 24 + lane * 24;
		uint2 a[DEPTH];
		unsigned c[DEPTH];
#pragma unroll
		for (int d = 0; d < DEPTH; ++d)
		{
			a[d] = __ldg(reinterpret_cast<const uint2*>(base + d * 768));
			c[d] = __ldg(reinterpret_cast<const unsigned*>(base + d * 768 + 8));
		}
#pragma unroll
		for (int s = 0; s < 10; ++s)
		{
			uint2 va = a[s % DEPTH];
			unsigned vc = c[s % DEPTH];
			if (s + DEPTH < 10)
			{
				a[s % DEPTH] = __ldg(reinterpret_cast<const uint2*>(base + (s + DEPTH) * 768));
				c[s % DEPTH] = __ldg(reinterpret_cast<const unsigned*>(base + (s + DEPTH) * 768 + 8));
			}
			// ~150 dependent-ish instructions of arithmetic per step
			float f = __uint_as_float((va.x & 0x007fffffu) | 0x3f800000u), g = __uint_as_float((va.y & 0x007fffffu) | 0x3f800000u);
#pragma unroll
			for (int k = 0; k < 70; ++k)
			{
				f = fmaf(f, g, 0.5f);
				g = fmaf(g, f, 0.25f);
			}
			acc += __float_as_uint(f) ^ __float_as_uint(g) ^ vc;
		}
	}
	if (acc == 0x12345678u)
		*out = acc;
}

__global__ void k_rec48(const char* __restrict__ p, size_t n, unsigned* out)
{
	unsigned acc = 0;
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
	{
		const uint4* r = reinterpret_cast<const uint4*>(p + i * 48);
		uint4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
		acc += a.x ^ b.y ^ c.z;
	}
	if (acc == 0x12345678u)
		*out = acc;
}

template <typename F>
static float timeit(F f, int reps = 20)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int i = 0; i < 3; ++i)
		f();
	cudaEventRecord(e0);
	for (int i = 0; i < reps; ++i)
		f();
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	return ms / reps;
}

int main()
{
	const size_t nrec = 10'000'000;
	const size_t bytes = nrec * 24; // 240 MB, > L2
	char *a, *b;
	unsigned *out, *ticket;
	cudaMalloc(&a, bytes);
	cudaMalloc(&b, bytes); // second buffer: alternate so that nothing is served by L2
	cudaMalloc(&out, 4);
	cudaMalloc(&ticket, 4);
	cudaMemset(a, 1, bytes);
	cudaMemset(b, 2, bytes);
	int sms = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int flip = 0;
	auto buf = [&]() { flip ^= 1; return flip ? a : b; };
	float ms;
	ms = timeit([&] { k_full<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(buf()), bytes / 16, out); });
	printf("a full 16B loads            %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { k_full<<<sms * 32, 256>>>(reinterpret_cast<const uint4*>(buf()), bytes / 16, out); });
	printf("a full 16B loads (x32 CTAs) %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { k_rec24<<<sms * 8, 256>>>(buf(), nrec, out); });
	printf("b 8+4 of 24B, grid-stride   %7.1f us  %6.0f GB/s (of 24 B/record)\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { k_rec24<<<(nrec + 255) / 256, 256>>>(buf(), nrec, out); });
	printf("b 8+4 of 24B, thread/record %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_rec24_persist<1><<<sms * 4, 256>>>(buf(), nrec, ticket, out); });
	printf("c persistent warps depth 1  %7.1f us  %6.0f GB/s (4 CTAs/SM, ~150 FMA per step)\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_rec24_persist<2><<<sms * 4, 256>>>(buf(), nrec, ticket, out); });
	printf("c persistent warps depth 2  %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_rec24_persist<4><<<sms * 4, 256>>>(buf(), nrec, ticket, out); });
	printf("c persistent warps depth 4  %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_rec24_persist<10><<<sms * 4, 256>>>(buf(), nrec, ticket, out); });
	printf("c persistent warps depth 10 %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_rec24_persist<2><<<sms * 8, 256>>>(buf(), nrec, ticket, out); });
	printf("c depth 2, 8 CTAs/SM        %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	ms = timeit([&] { k_rec48<<<(nrec / 2 + 255) / 256, 256>>>(buf(), nrec / 2, out); });
	printf("d 3x16 of 48B thread/record %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
