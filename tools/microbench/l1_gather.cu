// l1_gather.cu -- how does the L1TEX data pipe charge per-lane gathers?  (development microbenchmark)
// Each warp iteration fetches 32 "nodes" of 128 B (one per lane) from a table, in one of several ways:
//   0  per-lane: 4 x LDG.256 (lane reads its own 128-byte line)                 [what k_extend did]
//   1  cooperative: 4 x LDG.256, in each instruction 4 lanes read the 4 sectors of one line (8 lines / instr)
//   2  per-lane: 2 x LDG.256 (64-byte node)
//   3  per-lane: 1 x LDG.256
//   4  per-lane: 8 x LDG.128
//   5  cooperative: 1 x LDG.128 per line-quarter: 8 lanes read one line's 128 B as 8 x 16 B; 8 instr (4 lines / instr)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a l1_gather.cu -o l1_gather
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

struct __align__(32) F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg8(const void* p)
{
	F8 r;
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
	return r;
}
__device__ __forceinline__ float sum8(F8 v) { return v.a.x + v.a.y + v.a.z + v.a.w + v.b.x + v.b.y + v.b.z + v.b.w; }

template<int MODE>
__global__ void __launch_bounds__(128) k(const char* table, uint32_t lineMask, int iters, float* out)
{
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t state = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	float acc = 0.0f;
	for (int it = 0; it < iters; ++it)
	{
		state = state * 1664525u + 1013904223u;
		uint32_t line = (state >> 8) & lineMask;            // this lane's node
		// make the next address depend on loaded data a little (like traversal), without serialising everything
		if (MODE == 0)
		{
			const char* p = table + (size_t)line * 128;
			acc += sum8(ldg8(p)) + sum8(ldg8(p + 32)) + sum8(ldg8(p + 64)) + sum8(ldg8(p + 96));
		}
		else if (MODE == 1)
		{
			#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				const uint32_t l2 = __shfl_sync(0xFFFFFFFFu, line, (lane & ~3u) + j);
				acc += sum8(ldg8(table + (size_t)l2 * 128 + (lane & 3u) * 32));
			}
		}
		else if (MODE == 2)
		{
			const char* p = table + (size_t)line * 128;
			acc += sum8(ldg8(p)) + sum8(ldg8(p + 32));
		}
		else if (MODE == 3)
		{
			acc += sum8(ldg8(table + (size_t)line * 128));
		}
		else if (MODE == 4)
		{
			const float4* p = reinterpret_cast<const float4*>(table + (size_t)line * 128);
			#pragma unroll
			for (int j = 0; j < 8; ++j) { const float4 v = __ldg(p + j); acc += v.x + v.y + v.z + v.w; }
		}
		else if (MODE == 5)
		{
			#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				const uint32_t l2 = __shfl_sync(0xFFFFFFFFu, line, (lane & ~7u) + j);
				const float4 v = __ldg(reinterpret_cast<const float4*>(table + (size_t)l2 * 128) + (lane & 7u));
				acc += v.x + v.y + v.z + v.w;
			}
		}
		state += (acc > 1.0e30f) ? 1u : 0u;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template<int MODE>
static float run(const char* table, uint32_t lines, int iters, float* out, int grid)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	k<MODE><<<grid, 128>>>(table, lines - 1, iters / 8, out);
	cudaEventRecord(a);
	k<MODE><<<grid, 128>>>(table, lines - 1, iters, out);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	return ms;
}

int main()
{
	cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
	const int grid = prop.multiProcessorCount * 8;      // 32 warps per SM, like k_extend
	const int iters = 4096;
	float* out; cudaMalloc(&out, (size_t)grid * 128 * 4);
	const uint32_t sizes[] = { 512, 1u << 15, 1u << 20, 1u << 24 };   // lines: 64 KB (L1), 4 MB, 128 MB (~L2), 2 GB (DRAM)
	for (uint32_t lines : sizes)
	{
		char* table; cudaMalloc(&table, (size_t)lines * 128); cudaMemset(table, 0, (size_t)lines * 128);
		float t[6] = { run<0>(table, lines, iters, out, grid), run<1>(table, lines, iters, out, grid), run<2>(table, lines, iters, out, grid),
		               run<3>(table, lines, iters, out, grid), run<4>(table, lines, iters, out, grid), run<5>(table, lines, iters, out, grid) };
		const double warpIters = (double)grid * 4 * iters;
		printf("table %8.1f MB:", lines * 128.0 / 1e6);
		for (int m = 0; m < 6; ++m)
			printf("  mode%d %7.2f ms (%5.1f SM-cyc/warp-iter)", m, t[m], t[m] * 1e-3 * 1.9e9 * prop.multiProcessorCount / warpIters);
		printf("\n");
		cudaFree(table);
	}
	return 0;
}
