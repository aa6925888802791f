/*
 * gtb200/storage/b200.hpp -- storage traits for the B200 backend: `gridtools::storage::b200`.
 *
 *     auto field = gridtools::storage::builder<gridtools::storage::b200>.type<double>().dimensions(ni, nj, nk)
 *                      .halos(3, 3, 0).build();
 *
 * A storage traits type in GridTools is a tag with ADL friends (storage/traits.hpp:20-110; the reference's GPU traits:
 * storage/gpu.hpp:69-105).  This one has the device layout of storage::gpu -- i has stride 1, rows padded to 128
 * bytes, first interior element 128-byte aligned (what the TMA boxes of the named kernels and the coalesced rows of the
 * generic paths assume) -- and differs in how data_store::get_target_ptr() / host_view() move the data:
 *
 *   storage::gpu   one blocking cudaMemcpy of the whole padded allocation from / to the PAGEABLE host mirror (:86-99)
 *   storage::b200  gtb_staged_upload / gtb_staged_download: several host threads stage 8 MB chunks through a ring of
 *                  pinned buffers while the copy engine already moves the previous chunk; the upload returns as soon
 *                  as the host mirror has been read (the device side completes in stream order on the legacy stream,
 *                  like every kernel the reference launches), the download when the host mirror is complete.
 *
 * The host mirror itself cannot be pinned by a traits type: data_store allocates it with std::make_unique<T[]>
 * and frees it BEFORE the device holder (storage/data_store.hpp:101-104), so neither cudaHostAlloc nor
 * cudaHostRegister / cudaHostUnregister can be tied to its lifetime from here.  patches/gridtools-host-mirror-through-
 * traits.patch is the reference-side change that allows it: with it the mirror comes from
 * storage_allocate_host below and the transfers skip the staging ring.
 *
 * Plain host code over the C ABI (no CUDA headers needed); link with -lgtb200.
 */
#pragma once

#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>

#include <gridtools/common/integral_constant.hpp>
#include <gridtools/common/layout_map.hpp>
#include <gridtools/storage/gpu.hpp> // target_view (device-side element access of the views)

#include "../../gtb200.h"

namespace gridtools {
    namespace storage {
        namespace b200_impl_ {
            inline void check(int status, const char *what) {
                if (status != GTB_OK)
                    throw std::runtime_error(std::string(what) + ": " + gtb_last_error());
            }
            struct device_free {
                void operator()(void *p) const { gtb_device_free(p); }
            };
            struct host_free {
                void operator()(void *p) const { gtb_host_free(p); }
            };
        } // namespace b200_impl_

        struct b200 {
            friend std::false_type storage_is_host_referenceable(b200) { return {}; }

            template <size_t Dims>
            friend typename gpu_impl_::make_layout<Dims>::type storage_layout(b200, std::integral_constant<size_t, Dims>) {
                return {};
            }

            friend integral_constant<size_t, 128> storage_alignment(b200) { return {}; }

            template <class LazyType, class T = typename LazyType::type>
            friend auto storage_allocate(b200, LazyType, size_t size) {
                void *p = nullptr;
                b200_impl_::check(gtb_device_malloc(&p, (int64_t)(size * sizeof(T))), "gtb_device_malloc");
                return std::unique_ptr<T[], b200_impl_::device_free>(static_cast<T *>(p));
            }

            // The host mirror, page-locked -- used by a data_store that allocates its mirror through the traits
            // (patches/gridtools-host-mirror-through-traits.patch; the unpatched reference never asks).  The transfers
            // below then go from / to the mirror directly.
            template <class LazyType, class T = typename LazyType::type>
            friend auto storage_allocate_host(b200, LazyType, size_t size) {
                void *p = nullptr;
                b200_impl_::check(gtb_host_malloc(&p, (int64_t)(size * sizeof(T))), "gtb_host_malloc");
                return std::unique_ptr<T[], b200_impl_::host_free>(static_cast<T *>(p));
            }

            template <class T>
            friend void storage_update_target(b200, T *dst, T const *src, size_t size) {
                b200_impl_::check(gtb_staged_upload(const_cast<std::remove_cv_t<T> *>(dst),
                                      const_cast<std::remove_cv_t<T> *>(src), (int64_t)(size * sizeof(T)), nullptr),
                    "gtb_staged_upload");
            }

            template <class T>
            friend void storage_update_host(b200, T *dst, T const *src, size_t size) {
                b200_impl_::check(gtb_staged_download(const_cast<std::remove_cv_t<T> *>(dst),
                                      const_cast<std::remove_cv_t<T> *>(src), (int64_t)(size * sizeof(T)), nullptr),
                    "gtb_staged_download");
            }

            template <class T, class Info>
            friend gpu_impl_::target_view<T, Info> storage_make_target_view(b200, T *ptr, Info const &info) {
                return {ptr, info};
            }
        };
    } // namespace storage
} // namespace gridtools
