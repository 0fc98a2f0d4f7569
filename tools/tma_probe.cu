// tools/tma_probe.cu -- development probe: 3-D fp64 tensor-map box load with the wrappers of pft_kernels_tma.cuh
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#define PFT_STRICT_FP 1
#include "../porousfreezethaw_b200/csrc/pft_kernels_tma.cuh"
using namespace pft;

struct alignas(64) PP { CUtensorMap m; int bx, by, x, y, z; double *out; const CUtensorMap *gm; int use_global; };

__global__ void probe(const __grid_constant__ PP P)
{
	extern __shared__ __align__(128) double sm[];
	__shared__ uint64_t bar;
	if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	__syncthreads();
	if (threadIdx.x == 0) {
		mbar_arrive_expect_tx(&bar, P.bx * P.by * 8);
		tma_load_3d(sm, P.use_global ? P.gm : &P.m, P.x, P.y, P.z, &bar);
	}
	mbar_wait(&bar, 0);
	for (int i = threadIdx.x; i < P.bx * P.by; i += blockDim.x) P.out[i] = sm[i];
}

typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
	int use_global = argc > 1 ? atoi(argv[1]) : 0;
	int dtype = argc > 2 ? atoi(argv[2]) : 0;      // 0 f64, 1 u64, 2 u32 pairs
	int x0 = argc > 3 ? atoi(argv[3]) : -1;
	int y0 = argc > 6 ? atoi(argv[6]) : x0;
	int px = argc > 4 ? atoi(argv[4]) : 24, n2 = 20, nz = 34, bx = argc > 5 ? atoi(argv[5]) : 26, by = 22;
	std::vector<double> h((size_t)px * n2 * nz);
	for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
	double *d, *out; cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
	cudaMalloc(&out, bx * by * 8);
	void *fn; cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	PP P;
	cuuint64_t dim[3] = {(cuuint64_t)px, (cuuint64_t)n2, (cuuint64_t)nz}, str[2] = {(cuuint64_t)px * 8, (cuuint64_t)px * n2 * 8};
	cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
	CUtensorMapDataType dt = dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
	if (dtype == 2) { dim[0] *= 2; box[0] *= 2; }
	CUresult r = ((enc_t)fn)(&P.m, dt, 3, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode rc=%d\n", (int)r);
	CUtensorMap *gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &P.m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
	P.bx = bx; P.by = by; P.x = (dtype == 2 ? 2 * x0 : x0); P.y = y0; P.z = 1; P.out = out; P.gm = gm; P.use_global = use_global;
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
	probe<<<1, 128, bx * by * 8>>>(P);
	cudaError_t e = cudaDeviceSynchronize();
	printf("kernel: %s\n", cudaGetErrorString(e));
	std::vector<double> o(bx * by);
	cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
	int bad = 0;
	for (int y = 0; y < by; y++) for (int x = 0; x < bx; x++) {
		int gx = x + x0, gy = y + y0;
		double want = (gx < 0 || gy < 0 || gx >= px || gy >= n2) ? 0.0 : (double)((size_t)1 * px * n2 + (size_t)gy * px + gx);
		if (o[y * bx + x] != want) bad++;
	}
	printf("mismatches: %d  o[0]=%g o[bx+1]=%g\n", bad, o[0], o[bx + 1]);
	return 0;
}
