// tests/cuda_emu/cuda.h -- TEST INFRASTRUCTURE ONLY: the driver-API slice (tensor maps) that
// pft_b200.cu uses, for the CPU emulation build.  A CUtensorMap here is a plain description of
// the tensor and the box; pft_emu::tensor_load_3d copies a box with zero fill like the TMA unit.
#pragma once
#include <stdint.h>
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
typedef enum { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 10 } CUtensorMapDataType;
typedef enum { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 } CUtensorMapInterleave;
typedef enum { CU_TENSOR_MAP_SWIZZLE_NONE = 0 } CUtensorMapSwizzle;
typedef enum { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 } CUtensorMapL2promotion;
typedef enum { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 } CUtensorMapFloatOOBfill;
struct alignas(64) CUtensorMap {
	void *base;
	uint64_t dim[3];
	uint64_t stride[2];     // bytes
	uint32_t box[3];
	uint32_t pad_[9];
};
static_assert(sizeof(CUtensorMap) == 128, "emulated CUtensorMap keeps the real size");
CUresult cuTensorMapEncodeTiled(CUtensorMap *m, CUtensorMapDataType, cuuint32_t rank, void *addr, const cuuint64_t *dim,
                                const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *estr,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
