// tma.cuh -- host-side construction of TMA tensor maps over gtb_field layouts.
#pragma once

#include "common.cuh"

namespace gtb {

    template <class T>
    CUtensorMapDataType tma_dtype();
    template <>
    inline CUtensorMapDataType tma_dtype<double>() {
        return CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    }
    template <>
    inline CUtensorMapDataType tma_dtype<float>() {
        return CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    }

    // Tensor map whose coordinate 0 is element (-lead_i, -lead_j, 0) of the field and that covers len_i x len_j x
    // len_k elements from there.  Returns false when the layout is not TMA-addressable (base or strides not 16-byte
    // aligned): the caller then uses the cp.async variant.
    template <class T>
    bool make_map(CUtensorMap *map, const T *origin, int64_t sj, int64_t sk, int lead_i, int lead_j, int64_t len_i,
        int64_t len_j, int64_t len_k, int box_i, int box_j, int box_k = 1) {
        auto enc = tensor_map_encoder();
        if (!enc)
            return false;
        constexpr int es = sizeof(T);
        if ((sj * es) % 16 != 0 || (sk * es) % 16 != 0 || sj <= 0 || sk <= 0)
            return false;
        uintptr_t a = reinterpret_cast<uintptr_t>(origin - lead_i - (int64_t)lead_j * sj);
        if (a % 16 != 0)
            return false;
        cuuint64_t dims[3] = {(cuuint64_t)len_i, (cuuint64_t)len_j, (cuuint64_t)len_k};
        cuuint64_t strides[2] = {(cuuint64_t)(sj * es), (cuuint64_t)(sk * es)};
        cuuint32_t box[3] = {(cuuint32_t)box_i, (cuuint32_t)box_j, (cuuint32_t)box_k};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(map,
            tma_dtype<T>(),
            3,
            reinterpret_cast<void *>(a),
            dims,
            strides,
            box,
            estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS;
    }

} // namespace gtb
