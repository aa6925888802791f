/*
 * gtb200/stencil/b200_fused.hpp -- the fused generic path of the stencil::b200 backend: ONE launch per multi-stage for
 * specs that are not bound to a hand-written kernel (included by b200.hpp; needs nvcc).
 *
 * What it replaces in the reference's stencil::gpu backend, and how it differs:
 *
 *   a2  stage/multi-stage fusion (gpu/entry_point.hpp:97-124, be_api.hpp:211-222): all stages of a multi-stage run in
 *       one kernel, `__syncthreads` exactly where be_api's `need_sync` asks for it.
 *   a3  thread mapping (gpu/launch_kernel.hpp:46-166): the reference adds rows of 64 threads for the i-halo columns
 *       and masks them per cell.  Here a CTA has ONE thread per point of the halo-extended IJ tile,
 *       (BI + i-extent) x (BJ + j-extent) rounded up to whole warps, i fastest; a thread keeps its point for every
 *       cell of the multi-stage (so un-synchronised zero-offset dependencies stay inside a thread) and a cell masks
 *       the points outside its own extent.
 *   a4  ij caches (gpu/ij_cache.hpp:38-47, gpu/shared_allocator.hpp:21-53): shared-memory tiles with compile-time
 *       strides (1, tile width), built with sid::synthetic.
 *   a5  k loops (gpu/make_kernel_fun.hpp:33-142): parallel multi-stages take KB levels per CTA; forward / backward
 *       multi-stages sweep a column per thread.
 *   a6  k caches and their fill / flush (gpu/k_cache.hpp:24-75, gpu/fill_flush.hpp:123-324): register windows
 *       [kminus, kplus] slid once per level.  The reference rewrites the spec at compile time (extra fill and flush
 *       stages per elementary interval, bound checks chosen from the interval levels); here the loads and stores are
 *       part of the sweep itself and are checked at run time against the k bounds of the field: on the first level
 *       of the sweep the whole window is filled, afterwards the entry that slides in; every level flushes the entry
 *       that slides out, the last level the whole window.  Same values in the same places, no spec surgery.
 *   a7  temporaries that are neither ij- nor k-cached live in device memory: one whole-domain array when they are only
 *       read at IJ offset zero, CTA-private halo-extended blocks when they are read at IJ offsets
 *       (gpu/tmp_storage_sid.hpp:54-69), so that no CTA reads what another one writes.
 *   f4  horizontal multi-stages on REGISTER tiles (gpu_horizontal/entry_point.hpp:57-100, j_cache.hpp:31-52): a parallel
 *       multi-stage whose temporaries are all ij caches does not use a3/a4 at all -- one thread per column and level
 *       walks the rows of the block with the temporaries in per-thread register arrays that slide along j, the
 *       stages evaluated redundantly over their column extent, inputs with IJ offsets read from TMA-staged
 *       shared-memory tiles (hz_body; horizontal diffusion: 25 us against 47 us on shared-memory tiles).
 *   Not `fusable` (it takes the stage-by-stage path of b200.hpp): k caches inside a parallel multi-stage.  Sweeps
 *   with IJ extents work like the others (the halo threads sweep their columns with their own windows; a cache is
 *   filled / flushed on the columns inside its placeholder's IJ extent).  Consecutive column-local forward / backward
 *   multi-stages share ONE launch (body_chain).
 *
 * Everything here is written against a `Cta` policy (block / thread indices, barrier, shared-memory base): `cuda_cta`
 * is the product; tests/cpp/emulated_cta.hpp runs the very same body on the host, one OpenMP team per CTA, to pin
 * the index algebra against the reference's cpu_ifirst backend without a GPU (test infrastructure, not a fallback).
 */
#pragma once

#include <limits>
#include <map>
#include <memory>
#include <type_traits>
#include <vector>
#include <utility>

#include <gridtools/common/for_each.hpp>
#include <gridtools/common/functional.hpp>
#include <gridtools/common/host_device.hpp>
#include <gridtools/common/hymap.hpp>
#include <gridtools/common/integral_constant.hpp>
#include <gridtools/common/tuple.hpp>
#include <gridtools/common/tuple_util.hpp>
#include <gridtools/meta.hpp>
#include <gridtools/sid/allocator.hpp>
#include <gridtools/sid/as_const.hpp>
#include <gridtools/sid/block.hpp>
#include <gridtools/sid/composite.hpp>
#include <gridtools/sid/concept.hpp>
#include <gridtools/sid/contiguous.hpp>
#include <gridtools/sid/sid_shift_origin.hpp>
#include <gridtools/sid/synthetic.hpp>
#include <gridtools/stencil/be_api.hpp>
#include <gridtools/stencil/common/caches.hpp>
#include <gridtools/stencil/common/dim.hpp>
#include <gridtools/stencil/common/extent.hpp>

#ifdef __CUDACC__
#include <gridtools/common/cuda_util.hpp>
#endif

#include "../../gtb200.h"

namespace gridtools {
    namespace stencil {
        namespace b200_backend {
            namespace fused {
#ifdef __CUDACC__
                // ---------------------------------------------------------------- TMA (cp.async.bulk.tensor) helpers
                namespace tma {
                    __device__ __forceinline__ uint32_t smem_addr(const void *p) {
                        return static_cast<uint32_t>(__cvta_generic_to_shared(p));
                    }
                    __device__ __forceinline__ void mbar_init(uint64_t *bar) {
                        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
                        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                    }
                    __device__ __forceinline__ void expect_tx(uint64_t *bar, uint32_t bytes) {
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                                     : "memory");
                    }
                    __device__ __forceinline__ void load_3d(void *dst, const void *map, uint64_t *bar, int c0, int c1, int c2) {
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                                     "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_addr(dst)),
                                     "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
                                     : "memory");
                    }
                    __device__ __forceinline__ void wait(uint64_t *bar) { // phase 0: the barrier is used once per CTA
                        asm volatile("{\n"
                                     ".reg .pred p;\n"
                                     "WAIT_LOOP:\n"
                                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
                                     "@p bra WAIT_DONE;\n"
                                     "bra WAIT_LOOP;\n"
                                     "WAIT_DONE:\n"
                                     "}" ::"r"(smem_addr(bar))
                                     : "memory");
                    }
                } // namespace tma
#endif

                // ---------------------------------------------------------------- geometry
                // IJ block of a CTA, levels per CTA of a parallel multi-stage, unroll factor of the level loop of a sweep.
                // Defaults from profiles/r01_fused_generic.txt (256x256x80 fp64): hori_diff 32x8x8 53.0 us, 64x4x8
                // 61.5, 32x16x8 51.2, 80 levels per CTA 96.2; vert_adv chained, unroll 1 / 2 / 3: 168 / 130-151 / 128-132 us,
                // unroll 3 with prefetch 2 L1 / 4 L1 / 4 L2 / 8 L2: 108.8 / 111.9 / 112.1 / 156.0 us, sweeps in separate
                // launches 164-190 us.
                // ChainSweeps: consecutive forward / backward multi-stages in one launch (body_chain) -- what one of them
                // writes is then read with coherent loads by the next -- or one launch each, every field a multi-stage
                // only reads on the read-only path.
                // Prefetch > 0: a sweep asks for the lines of level k + Prefetch of every field it only reads (and of the
                // fields behind filled k caches) while it works on level k -- into L1 (PrefetchL1) or L2.  A sweep has one
                // thread per column, far fewer than an SM can hold, so it is bound by the latency of one level's loads;
                // the prefetch turns that into a bandwidth problem without touching the user functors.
                template <int_t BI = 32,
                    int_t BJ = 8,
                    int_t KB = 4, // (r02_fused_timing.txt: horizontal diffusion 46.8 us with 4 levels per CTA, 49.2 with 8, 61.6 with 16)
                    int_t SweepUnroll = 3,
                    bool ChainSweeps = true,
                    int_t Prefetch = 4,
                    bool PrefetchL1 = true,
                    int_t ParallelPrefetch = 0,
                    bool StageReadOnly = true,
                    bool RegisterTiles = true,
                    bool L2Hints = false>
                struct geometry {
                    static constexpr int_t bi = BI, bj = BJ, kb = KB, sweep_unroll = SweepUnroll, prefetch = Prefetch;
                    static constexpr bool chain_sweeps = ChainSweeps, prefetch_l1 = PrefetchL1;
                    // parallel multi-stages: the fields they only read (no k offsets) are STAGED -- one TMA box over the
                    // halo-extended IJ tile and the KB levels of the CTA lands in shared memory before the stages run, so
                    // that every stage argument with an IJ extent is a shared-memory read (see staged tiles below)
                    static constexpr bool stage_read_only = StageReadOnly;
                    // the same inside the KB levels a CTA of a parallel multi-stage walks (every thread asks for its
                    // own point of the halo-extended tile; not measured yet, hence off)
                    static constexpr int_t parallel_prefetch = ParallelPrefetch;
                    // parallel multi-stages whose temporaries are all ij caches: one thread per column i and level, the
                    // temporaries in per-thread REGISTER tiles that slide along j (no shared-memory tiles, no barriers
                    // between the stages; see hz_body)
                    static constexpr bool register_tiles = RegisterTiles;
                    // forward / backward multi-stages: what a sweep flushes into a temporary is stored with L2 priority
                    // `evict_last`, the temporaries and the fields a sweep only streams through are read `evict_first`,
                    // so that flush -> fill pairs between chained sweeps (vertical advection: ccol, dcol, 84 MB at
                    // 256x256x80) stay in the 126 MB L2 instead of going through HBM
                    static constexpr bool l2_hints = L2Hints;
                };

                template <class Extent>
                using has_ij_extent = std::bool_constant<Extent::iminus::value != 0 || Extent::iplus::value != 0 ||
                                                         Extent::jminus::value != 0 || Extent::jplus::value != 0>;

                // ---------------------------------------------------------------- what a placeholder is inside a MSS
                template <class Info>
                using is_ij_cached = std::is_same<typename Info::caches_t, meta::list<cache_type::ij>>;
                template <class Info>
                using is_k_cached = std::is_same<typename Info::caches_t, meta::list<cache_type::k>>;
                template <class Info>
                using is_plain = meta::is_empty<typename Info::caches_t>;
                template <class Info>
                using has_fill = meta::st_contains<typename Info::cache_io_policies_t, cache_io_policy::fill>;
                template <class Info>
                using has_flush = meta::st_contains<typename Info::cache_io_policies_t, cache_io_policy::flush>;
                template <class Info>
                using is_io_cached = std::bool_constant<is_k_cached<Info>::value &&
                                                        (has_fill<Info>::value || has_flush<Info>::value)>;
                // does the placeholder need memory behind it (device memory for a temporary)?
                template <class Info>
                using needs_memory = std::bool_constant<is_plain<Info>::value || is_io_cached<Info>::value>;

                /// Key of the field behind a filled / flushed k cache inside the composite of a multi-stage.
                template <class Plh>
                struct behind {};

                // ---------------------------------------------------------------- can a spec take the fused path?
                // the one shape that is not taken: k caches inside a parallel multi-stage (levels are spread over CTAs)
                template <class Mss>
                using mss_is_fusable = std::bool_constant<!be_api::is_parallel<typename Mss::execution_t>::value ||
                                                          !meta::any_of<is_k_cached, typename Mss::plh_map_t>::value>;

                template <class Spec, class Msses = be_api::make_fused_view<Spec>>
                using fusable = meta::all_of<mss_is_fusable, meta::rename<meta::list, Msses>>;

                // ---------------------------------------------------------------- shared-memory tiles (ij caches)
                template <class T, class Cta>
                struct tile_holder {
                    int_t m_bytes; // start of the tile in the CTA's dynamic shared memory (16-byte aligned)
                    int_t m_elems; // offset of the addressed element inside the tile
                    GT_FUNCTION T *operator()() const { return reinterpret_cast<T *>(Cta::smem() + m_bytes) + m_elems; }
                    friend GT_FUNCTION tile_holder operator+(tile_holder h, int_t d) {
                        h.m_elems += d;
                        return h;
                    }
                };

                template <class Tag>
                struct tile_kind {};

                // tile of a temporary with extent E under a BI x BJ block: (BI + i-extent) x (BJ + j-extent), i fastest,
                // origin at the block's first interior point
                template <class T, class Cta, class Geo, class Extent>
                auto make_tile(int_t &smem_bytes) {
                    constexpr int_t w = Geo::bi - Extent::iminus::value + Extent::iplus::value;
                    constexpr int_t h = Geo::bj - Extent::jminus::value + Extent::jplus::value;
                    const int_t start = (smem_bytes + 15) / 16 * 16;
                    smem_bytes = start + int_t(sizeof(T)) * w * h;
                    return sid::synthetic()
                        .template set<sid::property::origin>(
                            tile_holder<T, Cta>{start, -Extent::iminus::value - Extent::jminus::value * w})
                        .template set<sid::property::strides>(hymap::keys<dim::i, dim::j>::make_values(
                            integral_constant<int_t, 1>(), integral_constant<int_t, w>()))
                        .template set<sid::property::ptr_diff, int_t>()
                        .template set<sid::property::strides_kind, tile_kind<integral_constant<int_t, w>>>();
                }

                // ---------------------------------------------------------------- staged tiles (read-only fields)
                // The reference's GPU backend reads a field with an IJ extent through global loads at every access
                // (horizontal diffusion: 10 loads of `in` per point and level, behind each block barrier, SURVEY.md
                // section 8 row a4) and so did this path: a CTA walks its KB levels as a chain of barrier -> loads from
                // L1/L2 -> barrier, bound by their latency (52.9 us for horizontal diffusion at 256x256x80, 0.36 of the
                // roofline).  Staging: every plain field a parallel multi-stage only reads, without k offsets, gets a
                // shared-memory tile (BI + i-extent rounded up to 16 bytes) x (BJ + j-extent) x KB, filled ONCE per CTA
                // by one `cp.async.bulk.tensor.3d` (TMA) issued by thread 0 -- out-of-domain points come back as zeros
                // -- or, where the field is not TMA-addressable (and on the emulated CTAs of the tests), by a cooperative
                // copy.  The stages then see a SID whose strides are the tile's: nothing else changes for them.
                template <class T, class Cta, class Geo>
                struct staged_holder {
                    int_t m_bytes;   // start of the tile in the CTA's dynamic shared memory (128-byte aligned)
                    int_t m_elems;   // offset of the addressed element, including what the host shifted in k
                    int_t m_kstride; // elements per level of the tile
                    int_t m_k_first; // levels the host has already shifted by
                    GT_FUNCTION T const *operator()() const {
                        return reinterpret_cast<T const *>(Cta::smem() + m_bytes) +
                               (m_elems - (m_k_first + Cta::block_k() * Geo::kb) * m_kstride);
                    }
                    friend GT_FUNCTION staged_holder operator+(staged_holder h, int_t d) {
                        h.m_elems += d;
                        return h;
                    }
                };

                template <class T, class Geo, class Extent>
                struct staged_shape {
                    static constexpr int_t row = 16 / int_t(sizeof(T)) > 0 ? 16 / int_t(sizeof(T)) : 1;
                    static constexpr int_t used_w = Geo::bi - Extent::iminus::value + Extent::iplus::value;
                    static constexpr int_t w = (used_w + row - 1) / row * row; // box rows are multiples of 16 bytes
                    static constexpr int_t h = Geo::bj - Extent::jminus::value + Extent::jplus::value;
                    static constexpr int_t bytes = int_t(sizeof(T)) * w * h * Geo::kb;
                };

                template <class T, class Cta, class Geo, class Extent>
                auto make_staged_tile(int_t start, int_t k_first) {
                    using shape_t = staged_shape<T, Geo, Extent>;
                    return sid::synthetic()
                        .template set<sid::property::origin>(staged_holder<T, Cta, Geo>{
                            start, -Extent::iminus::value - Extent::jminus::value * shape_t::w, shape_t::w * shape_t::h, k_first})
                        .template set<sid::property::strides>(hymap::keys<dim::i, dim::j, dim::k>::make_values(
                            integral_constant<int_t, 1>(),
                            integral_constant<int_t, shape_t::w>(),
                            integral_constant<int_t, shape_t::w * shape_t::h>()))
                        .template set<sid::property::ptr_diff, int_t>()
                        .template set<sid::property::strides_kind,
                            tile_kind<meta::list<integral_constant<int_t, shape_t::w>, integral_constant<int_t, shape_t::h>>>>();
                }

                /// where a staged tile comes from: filled in by prepare_mss, read by the CTA that fills the tile
                template <class T>
                struct staged_src {
                    alignas(64) unsigned char m_map[128]; // CUtensorMap over the extended compute domain (if m_tma)
                    T const *m_origin;                    // element (0, 0, first level of the multi-stage)
                    ptrdiff_t m_sj, m_sk;
                    int_t m_start;                        // of the tile in shared memory
                    int_t m_tma;
                };

                template <class T>
                struct is_staged_type : std::bool_constant<std::is_same<T, double>::value || std::is_same<T, float>::value> {};

                // a raw pointer to T with i (unit), j and k strides -- and no further dimension -- behind the placeholder?
                template <class Key>
                using is_ijk_key = std::bool_constant<std::is_same<Key, dim::i>::value || std::is_same<Key, dim::j>::value ||
                                                      std::is_same<Key, dim::k>::value ||
                                                      std::is_same<Key, sid::blocked_dim<dim::i>>::value ||
                                                      std::is_same<Key, sid::blocked_dim<dim::j>>::value>;
                template <class Sid, class = void>
                struct is_stageable_sid : std::false_type {};
                template <class Sid>
                struct is_stageable_sid<Sid,
                    std::enable_if_t<std::is_pointer<sid::ptr_type<Sid>>::value &&
                                     meta::all_of<is_ijk_key, get_keys<sid::strides_type<Sid>>>::value &&
                                     has_key<sid::strides_type<Sid>, dim::i>::value &&
                                     has_key<sid::strides_type<Sid>, dim::j>::value &&
                                     has_key<sid::strides_type<Sid>, dim::k>::value>>
                    : std::bool_constant<std::is_same<std::decay_t<decltype(sid::get_stride<dim::i>(
                                                          std::declval<sid::strides_type<Sid> const &>()))>,
                          integral_constant<int_t, 1>>::value> {};

                // ---------------------------------------------------------------- register windows (k caches)
                template <class T, int_t Minus, int_t Plus>
                struct window {
                    static constexpr int_t minus = Minus, plus = Plus;
                    T m_v[Plus - Minus + 1];
                    GT_FUNCTION T *ptr() { return m_v - Minus; }
                    GT_FUNCTION void slide(integral_constant<int_t, 1>) {
#pragma unroll
                        for (int_t n = 0; n < Plus - Minus; ++n)
                            m_v[n] = m_v[n + 1];
                    }
                    GT_FUNCTION void slide(integral_constant<int_t, -1>) {
#pragma unroll
                        for (int_t n = Plus - Minus; n > 0; --n)
                            m_v[n] = m_v[n - 1];
                    }
                };
                template <class Info, class E = typename Info::extent_t>
                using window_of = window<std::remove_const_t<typename Info::data_t>, E::kminus::value, E::kplus::value>;

                // what stands for a k-cached placeholder inside the composite: no memory, k stride 1 (the window); the
                // real pointer is put in front of it per thread (hymap merge)
                template <class T>
                struct null_holder {
                    GT_FUNCTION T *operator()() const { return nullptr; }
                    friend GT_FUNCTION null_holder operator+(null_holder h, int_t) { return h; }
                };
                struct window_kind {};
                template <class T>
                auto make_window_stub() {
                    return sid::synthetic()
                        .template set<sid::property::origin>(null_holder<T>{})
                        .template set<sid::property::strides>(
                            hymap::keys<dim::k>::make_values(integral_constant<int_t, 1>()))
                        .template set<sid::property::ptr_diff, int_t>()
                        .template set<sid::property::strides_kind, window_kind>();
                }

                // ---------------------------------------------------------------- register tiles (ij caches, per thread)
                // What the reference's gpu_horizontal backend does with its j caches (gpu_horizontal/j_cache.hpp:31-52,
                // entry_point.hpp:57-100): a thread walks the rows of its block along j and keeps every temporary of the
                // multi-stage in a small register array that covers the temporary's IJ extent around the thread's column
                // and current row; the stages are evaluated for every column offset of their extent (redundantly, instead
                // of reading a neighbour's result through shared memory) and the arrays slide by one row per step.
                template <class T, class Extent>
                struct rtile {
                    static constexpr int_t iminus = Extent::iminus::value, jminus = Extent::jminus::value;
                    static constexpr int_t iw = Extent::iplus::value - iminus + 1, jw = Extent::jplus::value - jminus + 1;
                    T m_v[iw * jw];
                    GT_FUNCTION T *ptr() { return m_v + (-iminus * jw - jminus); }
                    GT_FUNCTION void slide() {
#pragma unroll
                        for (int_t i = 0; i < iw; ++i)
#pragma unroll
                            for (int_t j = 0; j + 1 < jw; ++j)
                                m_v[i * jw + j] = m_v[i * jw + j + 1];
                    }
                };
                template <class Info>
                using rtile_of = rtile<std::remove_const_t<typename Info::data_t>, typename Info::extent_t>;

                // what stands for such a temporary inside the composite: no memory, the strides of its register tile
                template <class JWidth>
                struct rtile_kind {};
                template <class T, class Extent>
                auto make_rtile_stub() {
                    using jw_t = integral_constant<int_t, rtile<T, Extent>::jw>;
                    return sid::synthetic()
                        .template set<sid::property::origin>(null_holder<T>{})
                        .template set<sid::property::strides>(
                            hymap::keys<dim::i, dim::j>::make_values(jw_t(), integral_constant<int_t, 1>()))
                        .template set<sid::property::ptr_diff, int_t>()
                        .template set<sid::property::strides_kind, rtile_kind<jw_t>>();
                }

                // can a multi-stage run on register tiles?  parallel, every temporary it touches is an ij cache
                template <class Info>
                using is_rtile_compatible = std::bool_constant<is_ij_cached<Info>::value || (is_plain<Info>::value && !Info::is_tmp_t::value)>;
                template <class Mss>
                using mss_is_horizontal = std::bool_constant<be_api::is_parallel<typename Mss::execution_t>::value &&
                                                             meta::all_of<is_rtile_compatible, typename Mss::plh_map_t>::value>;

                // Fields a multi-stage only reads go through the read-only data path (ld.global.nc), which also lets the
                // compiler move their loads across the stores of the sweep (the reference: gpu/entry_point.hpp:135-147).
                // Window and tile pointers are pointers to non-const and never take this overload.
                // L2 eviction priorities for the accesses of a sweep (geometry::l2_hints); 4- and 8-byte elements
                namespace l2 {
                    template <class T>
                    using hintable = std::bool_constant<std::is_arithmetic<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>;
#ifdef __CUDA_ARCH__
                    __device__ __forceinline__ uint64_t evict_first() {
                        uint64_t p;
                        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
                        return p;
                    }
                    __device__ __forceinline__ uint64_t evict_last() {
                        uint64_t p;
                        asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
                        return p;
                    }
                    template <bool NonCoherent, class T>
                    __device__ __forceinline__ T load_evict_first(T const *ptr) {
                        const uint64_t pol = evict_first();
                        T v;
                        if constexpr (sizeof(T) == 8) {
                            uint64_t r;
                            if constexpr (NonCoherent)
                                asm volatile("ld.global.nc.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(ptr), "l"(pol));
                            else
                                asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(ptr), "l"(pol) : "memory");
                            memcpy(&v, &r, 8);
                        } else {
                            uint32_t r;
                            if constexpr (NonCoherent)
                                asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(ptr), "l"(pol));
                            else
                                asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(ptr), "l"(pol) : "memory");
                            memcpy(&v, &r, 4);
                        }
                        return v;
                    }
                    template <class T>
                    __device__ __forceinline__ void store_evict_last(T *ptr, T v) {
                        const uint64_t pol = evict_last();
                        if constexpr (sizeof(T) == 8) {
                            uint64_t r;
                            memcpy(&r, &v, 8);
                            asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(ptr), "l"(r), "l"(pol) : "memory");
                        } else {
                            uint32_t r;
                            memcpy(&r, &v, 4);
                            asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(ptr), "r"(r), "l"(pol) : "memory");
                        }
                    }
#endif
                } // namespace l2

                // `Stream`: read-only fields are read `evict_first` (they pass through a sweep once); `DeadAfterRead`: keys
                // of temporaries an earlier sweep of the launch has written and this multi-stage only reads
                template <class ConstKeys, bool Stream = false, class DeadAfterRead = meta::list<>>
                struct read_only_deref {
                    template <class Key,
                        class T,
                        std::enable_if_t<meta::st_contains<ConstKeys, Key>::value && std::is_arithmetic<T>::value, int> = 0>
                    GT_FUNCTION T operator()(Key, T const *ptr) const {
#ifdef __CUDA_ARCH__
                        if constexpr (Stream && l2::hintable<T>::value)
                            return l2::load_evict_first<true>(ptr);
                        else
                            return __ldg(ptr);
#else
                        return *ptr;
#endif
                    }
                    template <class Key,
                        class T,
                        std::enable_if_t<!meta::st_contains<ConstKeys, Key>::value && meta::st_contains<DeadAfterRead, Key>::value &&
                                             l2::hintable<T>::value,
                            int> = 0>
                    GT_FUNCTION T operator()(Key, T const *ptr) const {
#ifdef __CUDA_ARCH__
                        return l2::load_evict_first<false>(ptr);
#else
                        return *ptr;
#endif
                    }
                    template <class Key, class Ptr>
                    GT_FUNCTION decltype(auto) operator()(Key, Ptr ptr) const {
                        return *ptr;
                    }
                };

                template <bool L1, class T>
                GT_FUNCTION void prefetch_line(T const *ptr) {
#ifdef __CUDA_ARCH__
                    if (L1)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
                    else
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
                    (void)ptr;
#endif
                }
                template <bool L1, class Other>
                GT_FUNCTION void prefetch_line(Other const &) {} // not a pointer to memory (global_parameter, ...)

                struct k_bounds {
                    int_t lo, hi; // valid levels of the field behind a cache, relative to the grid's k origin
                };

                // Fills the staged tiles of a CTA: levels [block_k * KB, block_k * KB + KB) of the halo-extended IJ tile of
                // every staged field.  TMA where the field is addressable (one elected thread, completion on an
                // mbarrier), a cooperative copy otherwise; points outside the extended compute domain are zeros.
                template <class Cta, class Geo, class StagedInfos, int_t Threads, class StagedSrcs>
                GT_FUNCTION void fill_staged_tiles(
                    int_t tid, StagedSrcs const &staged, int_t k_total, int_t ni, int_t nj, int_t bar_off) {
                    const int_t k0 = Cta::block_k() * Geo::kb, i0 = Cta::block_i() * Geo::bi, j0 = Cta::block_j() * Geo::bj;
#ifdef __CUDA_ARCH__
                    uint64_t *bar = reinterpret_cast<uint64_t *>(Cta::smem() + bar_off);
                    uint32_t tx_bytes = 0;
                    if (tid == 0) {
                        tma::mbar_init(bar);
                        tuple_util::host_device::for_each(
                            [&](auto info, auto const &src) GT_FORCE_INLINE_LAMBDA {
                                using info_t = decltype(info);
                                using shape_t = staged_shape<std::remove_const_t<typename info_t::data_t>, Geo,
                                    to_horizontal_extent<typename info_t::extent_t>>;
                                if (src.m_tma)
                                    tx_bytes += shape_t::bytes;
                            },
                            meta::rename<tuple, StagedInfos>(),
                            staged);
                        if (tx_bytes) {
                            tma::expect_tx(bar, tx_bytes);
                            tuple_util::host_device::for_each(
                                [&](auto, auto const &src) GT_FORCE_INLINE_LAMBDA {
                                    if (src.m_tma)
                                        tma::load_3d(Cta::smem() + src.m_start, src.m_map, bar, i0, j0, k0);
                                },
                                meta::rename<tuple, StagedInfos>(),
                                staged);
                        }
                    }
#endif
                    tuple_util::host_device::for_each(
                        [&](auto info, auto const &src) GT_FORCE_INLINE_LAMBDA {
                            using info_t = decltype(info);
                            using T = std::remove_const_t<typename info_t::data_t>;
                            using ext_t = to_horizontal_extent<typename info_t::extent_t>;
                            using shape_t = staged_shape<T, Geo, ext_t>;
                            if (src.m_tma)
                                return;
                            T *tile = reinterpret_cast<T *>(Cta::smem() + src.m_start);
                            for (int_t idx = tid; idx < shape_t::w * shape_t::h * Geo::kb; idx += Threads) {
                                const int_t x = idx % shape_t::w, r = idx / shape_t::w, y = r % shape_t::h, z = r / shape_t::h;
                                const int_t gi = i0 + ext_t::iminus::value + x, gj = j0 + ext_t::jminus::value + y,
                                            gk = k0 + z;
                                const bool inside = x < shape_t::used_w && gi < ni + ext_t::iplus::value &&
                                                    gj < nj + ext_t::jplus::value && gk < k_total;
                                tile[idx] = inside ? src.m_origin[gi + gj * src.m_sj + gk * src.m_sk] : T(0);
                            }
                        },
                        meta::rename<tuple, StagedInfos>(),
                        staged);
                    Cta::sync(); // the cooperative copies, and the mbarrier initialisation, are visible to every thread
#ifdef __CUDA_ARCH__
                    bool any_tma = false;
                    tuple_util::host_device::for_each(
                        [&](auto const &src) GT_FORCE_INLINE_LAMBDA { any_tma = any_tma || src.m_tma; }, staged);
                    if (any_tma)
                        tma::wait(bar);
#endif
                }

                // ---------------------------------------------------------------- per-thread body of one multi-stage
                // `Volatile`: placeholders written somewhere in the same launch (see body_chain), never read-only loaded
                template <class Cta,
                    class Mss,
                    class Geo,
                    class Volatile,
                    class Holder,
                    class Strides,
                    class KSizes,
                    class Bounds,
                    class StagedInfos = meta::list<>, // plh_infos of the staged fields ...
                    class StagedSrcs = tuple<>>       // ... and a staged_src for each of them
                struct mss_body {
                    using cta_t = Cta;
                    using extent_t = typename Mss::extent_t;
                    using plh_map_t = typename Mss::plh_map_t;
                    using k_cached_t = meta::filter<is_k_cached, plh_map_t>;
                    using io_cached_t = meta::filter<is_io_cached, plh_map_t>;
                    using step_t = typename Mss::k_step_t;
                    // (staged fields are shared-memory pointers: never through ld.global.nc)
                    template <class Info>
                    using is_read_only = std::bool_constant<Info::is_const_t::value &&
                                                            !meta::st_contains<Volatile, typename Info::plh_t>::value &&
                                                            !meta::st_contains<StagedInfos, Info>::value>;
                    // temporaries an earlier sweep of this launch wrote and this multi-stage only reads, uncached
                    template <class Info>
                    using is_dead_after_read = std::bool_constant<Info::is_const_t::value && Info::is_tmp_t::value &&
                                                                  is_plain<Info>::value &&
                                                                  meta::st_contains<Volatile, typename Info::plh_t>::value>;
                    static constexpr bool hints = Geo::l2_hints && !be_api::is_parallel<typename Mss::execution_t>::value;
                    using deref_t = read_only_deref<meta::transform<be_api::get_key, meta::filter<is_read_only, plh_map_t>>,
                        hints,
                        meta::if_c<hints,
                            meta::transform<be_api::get_key, meta::filter<is_dead_after_read, plh_map_t>>,
                            meta::list<>>>;

                    static constexpr bool parallel = be_api::is_parallel<typename Mss::execution_t>::value;
                    // column-local sweeps may share a launch with their neighbours (body_chain)
                    static constexpr bool chainable = !parallel && !has_ij_extent<extent_t>::value;
                    static constexpr int_t imin = extent_t::iminus::value, jmin = extent_t::jminus::value;
                    static constexpr int_t width = Geo::bi - imin + extent_t::iplus::value;
                    static constexpr int_t height = Geo::bj - jmin + extent_t::jplus::value;
                    static constexpr int_t threads = (width * height + 31) / 32 * 32;
                    static_assert(threads <= 1024, "stencil::b200 fused path: IJ block plus extents exceed one CTA");

                    Holder m_holder;   // composite of fields, tiles, window stubs; at the first level of the sweep
                    Strides m_strides; //
                    KSizes m_k_sizes;  // levels per elementary interval, in execution order
                    Bounds m_bounds;   // k_bounds per filled / flushed cache (order of io_cached_t)
                    int_t m_ni, m_nj;  // compute domain
                    int_t m_k_first;   // level of the first step of the sweep
                    StagedSrcs m_staged; // sources of the staged tiles
                    int_t m_bar;         // shared-memory offset of the mbarrier the TMA loads complete on

                    struct point {
                        int_t ti, tj, gi, gj;
                    };

                    template <class Extent>
                    GT_FUNCTION bool active(point const &p, Extent) const {
                        return p.ti >= Extent::iminus::value && p.ti < Geo::bi + Extent::iplus::value &&
                               p.tj >= Extent::jminus::value && p.tj < Geo::bj + Extent::jplus::value &&
                               p.gi < m_ni + Extent::iplus::value && p.gj < m_nj + Extent::jplus::value;
                    }

                    template <class Info, class Ptr>
                    GT_FUNCTION void exec_cells(Info, Ptr const &ptr, point const &p) const {
                        host_device::for_each<typename Info::cells_t>([&](auto cell) GT_FORCE_INLINE_LAMBDA {
                            if (decltype(cell.need_sync())::value)
                                Cta::sync();
                            if (active(p, cell.extent()))
                                cell.template operator()<deref_t>(ptr, m_strides);
                        });
                    }

                    GT_FUNCTION void operator()() const {
                        const int_t tid = Cta::tid();
                        point p;
                        p.ti = tid % width + imin;
                        p.tj = tid / width + jmin;
                        p.gi = Cta::block_i() * Geo::bi + p.ti;
                        p.gj = Cta::block_j() * Geo::bj + p.tj;
                        auto ptr = m_holder();
                        sid::shift(ptr, sid::get_stride<sid::blocked_dim<dim::i>>(m_strides), Cta::block_i());
                        sid::shift(ptr, sid::get_stride<sid::blocked_dim<dim::j>>(m_strides), Cta::block_j());
                        sid::shift(ptr, sid::get_stride<dim::i>(m_strides), p.ti);
                        sid::shift(ptr, sid::get_stride<dim::j>(m_strides), p.tj);
                        if constexpr (meta::length<StagedInfos>::value != 0)
                            fill_staged(tid);
                        run(std::bool_constant<parallel>(), ptr, p);
                    }

                    GT_FUNCTION void fill_staged(int_t tid) const {
                        int_t k_total = 0;
                        tuple_util::host_device::for_each([&](int_t size) GT_FORCE_INLINE_LAMBDA { k_total += size; }, m_k_sizes);
                        fill_staged_tiles<Cta, Geo, StagedInfos, threads>(tid, m_staged, k_total, m_ni, m_nj, m_bar);
                    }

                    // parallel multi-stage: this CTA takes levels [kb * KB, kb * KB + KB) of the multi-stage's interval
                    template <class Ptr>
                    GT_FUNCTION void run(std::true_type, Ptr &ptr, point const &p) const {
                        const int_t first = Cta::block_k() * Geo::kb, last = first + Geo::kb;
                        int_t cur = 0;
                        tuple_util::host_device::for_each(
                            [&](int_t size, auto info) GT_FORCE_INLINE_LAMBDA {
                                int_t lo = first - cur, hi = last - cur;
                                if (lo < 0)
                                    lo = 0;
                                if (hi > size)
                                    hi = size;
                                cur += size;
                                if (lo >= hi) {
                                    sid::shift(ptr, sid::get_stride<dim::k>(m_strides), size);
                                    return;
                                }
                                sid::shift(ptr, sid::get_stride<dim::k>(m_strides), lo);
                                for (int_t k = lo; k < hi; ++k) {
                                    prefetch_ahead<Geo::parallel_prefetch>(ptr, p, hi - k);
                                    exec_cells(info, ptr, p);
                                    info.inc_k(ptr, m_strides);
                                }
                                sid::shift(ptr, sid::get_stride<dim::k>(m_strides), size - hi);
                            },
                            m_k_sizes,
                            Mss::interval_infos());
                    }

                    // ---- k caches
                    template <class Info>
                    using behind_key = behind<typename Info::plh_t>;

                    // moves window entry `W` of one cache from (Fill) / to the field behind it if level k_pos + W exists there
                    template <bool Fill, class Info, int_t W, class Windows, class Ptr>
                    GT_FUNCTION void sync_entry(Windows &windows, Ptr const &ptr, int_t k_pos, k_bounds b) const {
                        if (k_pos + W < b.lo || k_pos + W >= b.hi)
                            return;
                        auto mem = host_device::at_key<behind_key<Info>>(ptr);
                        sid::shift(mem,
                            sid::get_stride_element<behind_key<Info>, dim::k>(m_strides),
                            integral_constant<int_t, W>());
                        auto &win = host_device::at_key<typename Info::key_t>(windows);
                        if constexpr (Fill) {
                            // a field that is only filled from (never flushed to, not written elsewhere in this
                            // launch) is read-only here: same non-coherent path as read_only_deref
                            using value_t = std::remove_cv_t<std::remove_reference_t<decltype(*mem)>>;
                            constexpr bool read_only = !has_flush<Info>::value && std::is_arithmetic<value_t>::value &&
                                                       !meta::st_contains<Volatile, typename Info::plh_t>::value;
#ifdef __CUDA_ARCH__
                            if constexpr (read_only && hints && l2::hintable<value_t>::value)
                                win.ptr()[W] = l2::load_evict_first<true>(&*mem);
                            else if constexpr (read_only)
                                win.ptr()[W] = __ldg(&*mem);
                            else
#endif
                                win.ptr()[W] = *mem;
                            (void)read_only;
                        } else {
                            using value_t = std::remove_cv_t<std::remove_reference_t<decltype(*mem)>>;
#ifdef __CUDA_ARCH__
                            // what a sweep flushes into a temporary is read again by a later sweep: keep it in L2
                            if constexpr (hints && Info::is_tmp_t::value && l2::hintable<value_t>::value)
                                l2::store_evict_last(&*mem, value_t(win.ptr()[W]));
                            else
#endif
                                *mem = win.ptr()[W];
                            (void)sizeof(value_t);
                        }
                    }

                    template <bool Fill, class Info, int_t From, int_t To, class Windows, class Ptr>
                    GT_FUNCTION void sync_range(Windows &windows, Ptr const &ptr, int_t k_pos, k_bounds b) const {
                        sync_entry<Fill, Info, From>(windows, ptr, k_pos, b);
                        if constexpr (From < To)
                            sync_range<Fill, Info, From + 1, To>(windows, ptr, k_pos, b);
                    }

                    // `whole`: first level of the sweep for fills, last level for flushes
                    template <bool Fill, class Windows, class Ptr>
                    GT_FUNCTION void sync_caches(
                        Windows &windows, Ptr const &ptr, point const &p, int_t k_pos, bool whole) const {
                        // the entry that enters (fill) or leaves (flush) the window at every step of the sweep
                        constexpr bool at_plus = (step_t::value > 0) == Fill;
                        tuple_util::host_device::for_each(
                            [&](auto info, k_bounds b) GT_FORCE_INLINE_LAMBDA {
                                using info_t = decltype(info);
                                using win_t = window_of<info_t>;
                                if constexpr (Fill ? has_fill<info_t>::value : has_flush<info_t>::value) {
                                    // only the columns on which the stages use this placeholder (its IJ extent)
                                    if (!active(p, to_horizontal_extent<typename info_t::extent_t>()))
                                        return;
                                    if (whole)
                                        sync_range<Fill, info_t, win_t::minus, win_t::plus>(windows, ptr, k_pos, b);
                                    else
                                        sync_entry<Fill, info_t, at_plus ? win_t::plus : win_t::minus>(
                                            windows, ptr, k_pos, b);
                                }
                            },
                            meta::rename<tuple, io_cached_t>(),
                            m_bounds);
                    }

                    // what a sweep reads from memory level by level: plain fields it does not write, fields behind fills
                    template <class Info>
                    using is_streamed_in = std::bool_constant<(is_plain<Info>::value && Info::is_const_t::value) ||
                                                              (is_k_cached<Info>::value && has_fill<Info>::value)>;
                    template <class Info>
                    using memory_key = meta::if_<is_plain<Info>, typename Info::key_t, behind<typename Info::plh_t>>;

                    template <int_t Distance, class Ptr>
                    GT_FUNCTION void prefetch_ahead(Ptr const &ptr, point const &p, int_t levels_left) const {
                        if constexpr (Distance > 0) {
                            if (levels_left <= Distance)
                                return; // level k + Distance is not part of this sweep / of this CTA's levels
                            host_device::for_each<meta::filter<is_streamed_in, plh_map_t>>([&](auto info)
                                                                                            GT_FORCE_INLINE_LAMBDA {
                                using info_t = decltype(info);
                                using key_t = memory_key<info_t>;
                                if (!active(p, to_horizontal_extent<typename info_t::extent_t>()))
                                    return;
                                auto mem = host_device::at_key<key_t>(ptr);
                                sid::shift(mem,
                                    sid::get_stride_element<key_t, dim::k>(m_strides),
                                    integral_constant<int_t, Distance * step_t::value>());
                                prefetch_line<Geo::prefetch_l1>(mem);
                            });
                        }
                    }

                    // forward / backward multi-stage: one column per thread -- the threads of the halo points sweep their
                    // columns too, with their own windows -- k caches in registers
                    template <class Ptr>
                    GT_FUNCTION void run(std::false_type, Ptr &ptr, point const &p) const {
                        using keys_t = meta::transform<be_api::get_key, k_cached_t>;
                        using windows_t = hymap::from_keys_values<keys_t, meta::transform<window_of, k_cached_t>>;
                        windows_t windows;
                        auto mixed = hymap::host_device::merge(
                            tuple_util::host_device::transform(
                                [](auto &w) GT_FORCE_INLINE_LAMBDA { return w.ptr(); }, windows),
                            std::move(ptr));
                        int_t total = 0;
                        tuple_util::host_device::for_each(
                            [&](int_t size) GT_FORCE_INLINE_LAMBDA { total += size; }, m_k_sizes);
                        int_t n = 0, k_pos = m_k_first;
                        tuple_util::host_device::for_each(
                            [&](int_t size, auto info) GT_FORCE_INLINE_LAMBDA {
#pragma unroll(Geo::sweep_unroll)
                                for (int_t k = 0; k < size; ++k) {
                                    prefetch_ahead<Geo::prefetch>(mixed.secondary(), p, total - n);
                                    sync_caches<true>(windows, mixed.secondary(), p, k_pos, n == 0);
                                    exec_cells(info, mixed, p);
                                    sync_caches<false>(windows, mixed.secondary(), p, k_pos, n == total - 1);
                                    tuple_util::host_device::for_each(
                                        [](auto &w) GT_FORCE_INLINE_LAMBDA { w.slide(step_t()); }, windows);
                                    info.inc_k(mixed.secondary(), m_strides);
                                    k_pos += step_t::value;
                                    ++n;
                                }
                            },
                            m_k_sizes,
                            Mss::interval_infos());
                    }
                };

                // ---------------------------------------------------------------- parallel multi-stage on register tiles
                // One thread per column i of the block and per level of the CTA's KB levels (BI x KB threads); the thread
                // walks the BJ rows of the block, starting as many rows early as the temporaries need to be warm.  Fields
                // read with IJ offsets come from the staged shared-memory tiles (TMA) like on the tile path; nothing is
                // exchanged between threads, so there is no barrier after the tiles have landed.
                template <class Cta,
                    class Mss,
                    class Geo,
                    class Holder,
                    class Strides,
                    class KSizes,
                    class StagedInfos = meta::list<>,
                    class StagedSrcs = tuple<>>
                struct hz_body {
                    using cta_t = Cta;
                    using register_tiles_t = void; // (how the tests tell this body from mss_body)
                    using extent_t = typename Mss::extent_t;
                    using plh_map_t = typename Mss::plh_map_t;
                    using cached_t = meta::filter<is_ij_cached, plh_map_t>;
                    template <class Info>
                    using is_read_only =
                        std::bool_constant<Info::is_const_t::value && !meta::st_contains<StagedInfos, Info>::value>;
                    using deref_t = read_only_deref<meta::transform<be_api::get_key, meta::filter<is_read_only, plh_map_t>>>;

                    static constexpr bool parallel = true, chainable = false;
                    static constexpr int_t threads = (Geo::bi * Geo::kb + 31) / 32 * 32;
                    static_assert(threads <= 1024, "stencil::b200 register-tile path: BI x KB exceeds one CTA");
                    // first row of the walk, relative to the block: the stages run `jplus` rows ahead of their consumers
                    static constexpr int_t j_start = extent_t::jminus::value - extent_t::jplus::value;

                    Holder m_holder;
                    Strides m_strides;
                    KSizes m_k_sizes;
                    int_t m_ni, m_nj;
                    StagedSrcs m_staged;
                    int_t m_bar;

                    // one cell at column offsets [iminus, iplus] of its extent, `jplus` rows ahead of row j
                    template <int_t J, class Cell, class Mixed>
                    GT_FUNCTION void exec_cell(Cell cell, Mixed const &mixed, int_t rows) const {
                        using ext_t = typename Cell::extent_t;
                        constexpr int_t jj = J + ext_t::jplus::value; // the row this cell computes in step J
                        if constexpr (jj >= ext_t::jminus::value) {
                            if (jj < rows + ext_t::jplus::value) {
#pragma unroll
                                for (int_t di = ext_t::iminus::value; di <= ext_t::iplus::value; ++di) {
                                    Mixed q = mixed;
                                    tuple_util::host_device::for_each(
                                        [di](auto &wp, auto info) GT_FORCE_INLINE_LAMBDA {
                                            wp += di * rtile_of<decltype(info)>::jw + ext_t::jplus::value;
                                        },
                                        q.primary(),
                                        meta::rename<tuple, cached_t>());
                                    sid::shift(q.secondary(), sid::get_stride<dim::i>(m_strides), di);
                                    sid::shift(q.secondary(), sid::get_stride<dim::j>(m_strides), typename ext_t::jplus());
                                    cell.template operator()<deref_t>(q, m_strides);
                                }
                            }
                        }
                    }

                    template <int_t J, class Info, class Tiles, class Mixed>
                    GT_FUNCTION void walk(Info info, Tiles &tiles, Mixed &mixed, int_t rows) const {
                        if constexpr (J < Geo::bj) {
                            if (J < rows) {
                                host_device::for_each<typename Info::cells_t>(
                                    [&](auto cell) GT_FORCE_INLINE_LAMBDA { exec_cell<J>(cell, mixed, rows); });
                                sid::shift(mixed.secondary(), sid::get_stride<dim::j>(m_strides), integral_constant<int_t, 1>());
                                tuple_util::host_device::for_each([](auto &t) GT_FORCE_INLINE_LAMBDA { t.slide(); }, tiles);
                                walk<J + 1>(info, tiles, mixed, rows);
                            }
                        }
                    }

                    GT_FUNCTION void operator()() const {
                        const int_t tid = Cta::tid();
                        int_t k_total = 0;
                        tuple_util::host_device::for_each([&](int_t size) GT_FORCE_INLINE_LAMBDA { k_total += size; }, m_k_sizes);
                        if constexpr (meta::length<StagedInfos>::value != 0)
                            fill_staged_tiles<Cta, Geo, StagedInfos, threads>(tid, m_staged, k_total, m_ni, m_nj, m_bar);
                        const int_t ti = tid % Geo::bi, level = Cta::block_k() * Geo::kb + tid / Geo::bi;
                        const int_t gi = Cta::block_i() * Geo::bi + ti;
                        int_t rows = m_nj - Cta::block_j() * Geo::bj; // rows of this block inside the compute domain
                        if (rows > Geo::bj)
                            rows = Geo::bj;
                        if (tid >= Geo::bi * Geo::kb || gi >= m_ni || level >= k_total)
                            return;
                        auto ptr = m_holder();
                        sid::shift(ptr, sid::get_stride<sid::blocked_dim<dim::i>>(m_strides), Cta::block_i());
                        sid::shift(ptr, sid::get_stride<sid::blocked_dim<dim::j>>(m_strides), Cta::block_j());
                        sid::shift(ptr, sid::get_stride<dim::i>(m_strides), ti);
                        sid::shift(ptr, sid::get_stride<dim::j>(m_strides), integral_constant<int_t, j_start>());
                        sid::shift(ptr, sid::get_stride<dim::k>(m_strides), level);
                        int_t cur = 0;
                        tuple_util::host_device::for_each(
                            [&](int_t size, auto info) GT_FORCE_INLINE_LAMBDA {
                                const bool mine = level >= cur && level < cur + size;
                                cur += size;
                                if (!mine)
                                    return;
                                using keys_t = meta::transform<be_api::get_key, cached_t>;
                                using tiles_t = hymap::from_keys_values<keys_t, meta::transform<rtile_of, cached_t>>;
                                tiles_t tiles;
                                auto mixed = hymap::host_device::merge(
                                    tuple_util::host_device::transform(
                                        [](auto &t) GT_FORCE_INLINE_LAMBDA { return t.ptr(); }, tiles),
                                    ptr);
                                walk<j_start>(info, tiles, mixed, rows);
                            },
                            m_k_sizes,
                            Mss::interval_infos());
                    }
                };

                // ---------------------------------------------------------------- consecutive sweeps in one launch
                // Forward / backward multi-stages that follow each other are column-local on both sides (no IJ extents,
                // one thread per column), so they run back to back inside one kernel: what the first leaves in device
                // memory (flushed k caches, plain temporaries) is read again by the thread that wrote it.  This is the
                // case the reference fuses with launch_or_fuse (gpu/entry_point.hpp:97-124); vertical advection and the
                // Thomas solve become one launch.
                template <class... Bodies>
                struct body_chain {
                    using first_t = meta::first<body_chain>;
                    using cta_t = typename first_t::cta_t;
                    static constexpr bool parallel = false, chainable = true;
                    static constexpr int_t threads = first_t::threads;
                    static_assert(std::conjunction<std::bool_constant<Bodies::threads == threads>...>::value &&
                                      std::conjunction<std::bool_constant<Bodies::chainable>...>::value,
                        GT_INTERNAL_ERROR);
                    tuple<Bodies...> m_bodies;

                    GT_FUNCTION void operator()() const {
                        tuple_util::host_device::for_each(
                            [](auto const &body) GT_FORCE_INLINE_LAMBDA {
                                body();
                                cta_t::sync(); // the next multi-stage may reuse the shared-memory tiles
                            },
                            m_bodies);
                    }
                };

                template <class... Bodies, class Body>
                body_chain<Bodies..., Body> chained(body_chain<Bodies...> chain, Body body) {
                    return {tuple_util::push_back(std::move(chain.m_bodies), std::move(body))};
                }
                template <class First, class Body, std::enable_if_t<!meta::is_instantiation_of<body_chain, First>::value, int> = 0>
                body_chain<First, Body> chained(First first, Body body) {
                    return {tuple<First, Body>{std::move(first), std::move(body)}};
                }

                template <class Body>
                struct pending {
                    Body m_body;
                    int_t m_nbi, m_nbj, m_nbk, m_smem;

                    template <class Launcher>
                    void flush(Launcher &launcher) && {
                        if (m_nbi > 0 && m_nbj > 0 && m_nbk > 0)
                            launcher.launch(m_body, m_nbi, m_nbj, m_nbk, Body::threads, m_smem);
                    }
                    template <bool Chain, class Launcher, class Other>
                    auto then(Launcher &launcher, pending<Other> next) && {
                        if constexpr (Chain && Body::chainable && Other::chainable) {
                            auto chain = chained(std::move(m_body), std::move(next.m_body));
                            return pending<decltype(chain)>{std::move(chain),
                                m_nbi,
                                m_nbj,
                                m_nbk > next.m_nbk ? m_nbk : next.m_nbk,
                                m_smem > next.m_smem ? m_smem : next.m_smem};
                        } else {
                            std::move(*this).flush(launcher);
                            return next;
                        }
                    }
                };
                struct nothing_pending {
                    template <class Launcher>
                    void flush(Launcher &) && {}
                    template <bool Chain, class Launcher, class Next>
                    Next then(Launcher &, Next next) && {
                        return next;
                    }
                };

                // ---------------------------------------------------------------- host side: one multi-stage
                template <class Plh, class DataStores>
                k_bounds field_k_bounds(DataStores const &data_stores) {
                    auto const &store = at_key<Plh>(data_stores);
                    return {int_t(sid::get_lower_bound<dim::k>(sid::get_lower_bounds(store))),
                        int_t(sid::get_upper_bound<dim::k>(sid::get_upper_bounds(store)))};
                }

                // which placeholders of a multi-stage are staged (see staged tiles above)
                template <class Mss, class Geo, class DataStores>
                struct staged_in {
                    template <class Info, class Plh = typename Info::plh_t>
                    using store_t = std::decay_t<decltype(at_key<Plh>(std::declval<DataStores &>()))>;
                    template <class Info>
                    using apply = std::bool_constant<Geo::stage_read_only &&
                        be_api::is_parallel<typename Mss::execution_t>::value && is_plain<Info>::value &&
                        Info::is_const_t::value && !Info::is_tmp_t::value && Info::extent_t::kminus::value == 0 &&
                        Info::extent_t::kplus::value == 0 && is_staged_type<std::remove_const_t<typename Info::data_t>>::value &&
                        is_stageable_sid<store_t<Info>>::value>;
                };
                template <class Mss, class Grid>
                int_t k_total_of(Mss, Grid const &grid) {
                    int_t n = 0;
                    tuple_util::for_each([&](int_t size) { n += size; }, be_api::make_k_sizes(Mss::interval_infos(), grid));
                    return n;
                }

                template <class Launcher, class Geo, class Volatile, class Mss, class Grid, class DataStores>
                auto prepare_mss(Mss, Grid const &grid, DataStores &data_stores) {
                    using cta_t = typename Launcher::cta_t;
                    using plh_map_t = typename Mss::plh_map_t;
                    using io_cached_t = meta::filter<is_io_cached, plh_map_t>;
                    int_t smem_bytes = 0;
                    // (a geometry with more than 1024 column-levels per CTA keeps the tile path)
                    constexpr bool horizontal =
                        Geo::register_tiles && mss_is_horizontal<Mss>::value && Geo::bi * Geo::kb <= 1024;
                    const int_t k_first = grid.k_start(Mss::interval(), Mss::execution());
                    // plain fields a parallel multi-stage only reads, without k offsets: staged through shared memory
                    using staged_t = meta::filter<staged_in<Mss, Geo, DataStores>::template apply, plh_map_t>;
                    int_t staged_starts[meta::length<staged_t>::value + 1] = {}, staged_count = 0;

                    // fields, shared-memory tiles and window stubs under the keys the stages look them up with ...
                    auto members = tuple_util::transform(
                        overload(
                            [&](meta::list<cache_type::ij>, auto info) {
                                using info_t = decltype(info);
                                using data_t = std::remove_const_t<typename info_t::data_t>;
                                if constexpr (horizontal)
                                    return make_rtile_stub<data_t, typename info_t::extent_t>();
                                else
                                    return make_tile<data_t, cta_t, Geo, typename info_t::extent_t>(smem_bytes);
                            },
                            [](meta::list<cache_type::k>, auto info) {
                                return make_window_stub<std::remove_const_t<typename decltype(info)::data_t>>();
                            },
                            [&](meta::list<>, auto info) {
                                using info_t = decltype(info);
                                if constexpr (meta::st_contains<staged_t, info_t>::value) {
                                    using shape_t = staged_shape<std::remove_const_t<typename info_t::data_t>, Geo,
                                        to_horizontal_extent<typename info_t::extent_t>>;
                                    const int_t start = (smem_bytes + 127) / 128 * 128;
                                    smem_bytes = start + shape_t::bytes;
                                    staged_starts[staged_count++] = start; // (transform visits the placeholders in order)
                                    return make_staged_tile<std::remove_const_t<typename info_t::data_t>, cta_t, Geo,
                                        to_horizontal_extent<typename info_t::extent_t>>(start, k_first);
                                } else {
                                    return sid::add_const(info.is_const(), at_key<decltype(info.plh())>(data_stores));
                                }
                            }),
                        meta::rename<tuple, meta::transform<be_api::get_caches, plh_map_t>>(),
                        meta::rename<tuple, plh_map_t>());
                    // where the staged tiles come from (same walk through the shared memory as above)
                    int_t staged_walk = 0;
                    auto staged_srcs = tuple_util::transform(
                        [&](auto info) {
                            using info_t = decltype(info);
                            using T = std::remove_const_t<typename info_t::data_t>;
                            using ext_t = to_horizontal_extent<typename info_t::extent_t>;
                            using shape_t = staged_shape<T, Geo, ext_t>;
                            auto &store = at_key<decltype(info.plh())>(data_stores);
                            auto store_strides = sid::get_strides(store);
                            staged_src<T> src{};
                            src.m_sj = sid::get_stride<dim::j>(store_strides);
                            src.m_sk = sid::get_stride<dim::k>(store_strides);
                            src.m_origin = sid::get_origin(store)() + ptrdiff_t(k_first) * src.m_sk;
                            src.m_start = staged_starts[staged_walk++];
                            // the tensor is the extended compute domain over the levels of this multi-stage: what
                            // lies outside reads as zero, like in the cooperative copy
                            const int64_t dims[3] = {grid.i_size() - ext_t::iminus::value + ext_t::iplus::value,
                                grid.j_size() - ext_t::jminus::value + ext_t::jplus::value,
                                k_total_of(Mss(), grid)};
                            const int64_t byte_strides[2] = {int64_t(src.m_sj) * int64_t(sizeof(T)), int64_t(src.m_sk) * int64_t(sizeof(T))};
                            const int box[3] = {shape_t::w, shape_t::h, Geo::kb};
                            src.m_tma = dims[0] > 0 && dims[1] > 0 && dims[2] > 0 &&
                                        Launcher::tensor_map(src.m_map,
                                            src.m_origin + ext_t::iminus::value + ptrdiff_t(ext_t::jminus::value) * src.m_sj,
                                            int(sizeof(T)), dims, byte_strides, box);
                            return src;
                        },
                        meta::rename<tuple, staged_t>());
                    const int_t bar_off = (smem_bytes + 7) / 8 * 8;
                    if (meta::length<staged_t>::value != 0)
                        smem_bytes = bar_off + 8;
                    // ... plus the fields behind filled / flushed k caches
                    auto behinds = tuple_util::transform(
                        [&](auto info) {
                            return sid::add_const(std::bool_constant<!has_flush<decltype(info)>::value>(),
                                at_key<decltype(info.plh())>(data_stores));
                        },
                        meta::rename<tuple, io_cached_t>());
                    using keys_t = meta::rename<sid::composite::keys,
                        meta::concat<meta::transform<be_api::get_key, plh_map_t>,
                            meta::transform<behind, meta::transform<be_api::get_plh, io_cached_t>>>>;
                    auto composite = tuple_util::convert_to<keys_t::template values>(
                        tuple_util::concat(std::move(members), std::move(behinds)));

                    auto bounds = tuple_util::transform(
                        [&](auto info) { return field_k_bounds<decltype(info.plh())>(data_stores); },
                        meta::rename<tuple, io_cached_t>());

                    auto strides = sid::get_strides(composite);
                    sid::ptr_diff_type<decltype(composite)> offset{};
                    sid::shift(offset, sid::get_stride<dim::k>(strides), k_first);
                    auto k_sizes = be_api::make_k_sizes(Mss::interval_infos(), grid);
                    int_t k_total = 0;
                    tuple_util::for_each([&](int_t n) { k_total += n; }, k_sizes);
                    const int_t ni = grid.i_size(), nj = grid.j_size();
                    const int_t nbi = ni > 0 ? (ni + Geo::bi - 1) / Geo::bi : 0, nbj = nj > 0 ? (nj + Geo::bj - 1) / Geo::bj : 0;
                    if constexpr (horizontal) {
                        using body_t = hz_body<cta_t,
                            Mss,
                            Geo,
                            decltype(sid::get_origin(composite) + offset),
                            decltype(strides),
                            decltype(k_sizes),
                            staged_t,
                            decltype(staged_srcs)>;
                        return pending<body_t>{
                            body_t{sid::get_origin(composite) + offset, strides, k_sizes, ni, nj, staged_srcs, bar_off},
                            nbi,
                            nbj,
                            (k_total + Geo::kb - 1) / Geo::kb,
                            smem_bytes};
                    } else {
                        using body_t = mss_body<cta_t,
                            Mss,
                            Geo,
                            meta::if_<be_api::is_parallel<typename Mss::execution_t>, meta::list<>, Volatile>,
                            decltype(sid::get_origin(composite) + offset),
                            decltype(strides),
                            decltype(k_sizes),
                            decltype(bounds),
                            staged_t,
                            decltype(staged_srcs)>;
                        const int_t nbk = body_t::parallel ? (k_total + Geo::kb - 1) / Geo::kb : (k_total > 0 ? 1 : 0);
                        return pending<body_t>{
                            body_t{sid::get_origin(composite) + offset, strides, k_sizes, bounds, ni, nj, k_first, staged_srcs, bar_off},
                            nbi,
                            nbj,
                            nbk,
                            smem_bytes};
                    }
                }

                template <class Launcher, class Geo, class Volatile, class Grid, class DataStores, class Pending>
                void launch_msses(Launcher &launcher, meta::list<>, Grid const &, DataStores &, Pending pending) {
                    std::move(pending).flush(launcher);
                }
                template <class Launcher,
                    class Geo,
                    class Volatile,
                    class Mss,
                    class... Msses,
                    class Grid,
                    class DataStores,
                    class Pending>
                void launch_msses(Launcher &launcher,
                    meta::list<Mss, Msses...>,
                    Grid const &grid,
                    DataStores &data_stores,
                    Pending pending) {
                    launch_msses<Launcher, Geo, Volatile>(launcher,
                        meta::list<Msses...>(),
                        grid,
                        data_stores,
                        std::move(pending).template then<Geo::chain_sweeps>(
                            launcher, prepare_mss<Launcher, Geo, Volatile>(Mss(), grid, data_stores)));
                }

                // placeholders some sweep of the spec writes: sweeps may share a launch (body_chain), and what one of
                // them writes must not be read through the non-coherent path by the next
                template <class Mss>
                using written_plhs = meta::transform<be_api::get_plh,
                    meta::filter<meta::not_<be_api::get_is_const>::apply, typename Mss::plh_map_t>>;
                template <class Mss>
                using is_sweep = std::bool_constant<!be_api::is_parallel<typename Mss::execution_t>::value>;
                template <class Msses>
                using written_in_sweeps = meta::dedup<meta::rename<meta::concat,
                    meta::push_front<meta::transform<written_plhs, meta::filter<is_sweep, meta::rename<meta::list, Msses>>>,
                        meta::list<>>>>;

                // ---------------------------------------------------------------- temporaries in device memory
                // read at offset zero only: one array over the whole domain, i fastest like the fields
                template <class Geo, class T, class Extent, class Grid, class Interval, class Allocator>
                auto make_temporary(std::false_type, Extent extent, Grid const &grid, Interval interval, Allocator &alloc) {
                    auto offsets = hymap::keys<dim::k>::make_values(-grid.k_start(interval) - extent.minus(dim::k()));
                    // first key = stride 1 (stride_util::make_strides_from_sizes)
                    auto sizes = hymap::keys<dim::i, dim::j, dim::k>::make_values(
                        grid.i_size(), grid.j_size(), grid.k_size(interval, extent));
                    using kind_t = meta::list<Extent, behind<void>>;
                    return sid::block(
                        sid::shift_sid_origin(sid::make_contiguous<T, ptrdiff_t, kind_t>(alloc, sizes), offsets),
                        hymap::keys<dim::i, dim::j>::make_values(
                            integral_constant<int_t, Geo::bi>(), integral_constant<int_t, Geo::bj>()));
                }
                // read at IJ offsets: every CTA owns a halo-extended copy of its block (the stages that write it run
                // on the extended tile), so no CTA ever reads what another one writes -- the layout idea of the
                // reference's gpu/tmp_storage_sid.hpp:54-69 with the block index outermost but k
                template <class Geo, class T, class Extent, class Grid, class Interval, class Allocator>
                auto make_temporary(std::true_type, Extent extent, Grid const &grid, Interval interval, Allocator &alloc) {
                    auto offsets = hymap::keys<dim::i, dim::j, dim::k>::make_values(-extent.minus(dim::i()),
                        -extent.minus(dim::j()),
                        -grid.k_start(interval) - extent.minus(dim::k()));
                    auto sizes = hymap::
                        keys<dim::i, dim::j, sid::blocked_dim<dim::i>, sid::blocked_dim<dim::j>, dim::k>::make_values(
                            extent.extend(dim::i(), integral_constant<int_t, Geo::bi>()),
                            extent.extend(dim::j(), integral_constant<int_t, Geo::bj>()),
                            (grid.i_size() + Geo::bi - 1) / Geo::bi,
                            (grid.j_size() + Geo::bj - 1) / Geo::bj,
                            grid.k_size(interval, extent));
                    using kind_t = meta::list<Extent, behind<Geo>>;
                    return sid::shift_sid_origin(sid::make_contiguous<T, ptrdiff_t, kind_t>(alloc, sizes), offsets);
                }

                // ---------------------------------------------------------------- host side: the whole spec
                // (not called `run`: the tag's geometry argument makes this namespace an associated one of every b200<> tag,
                // and an unqualified run(spec, backend, grid, ...) with an lvalue spec would pick an overload named run here)
                template <class Geo = geometry<>, class Launcher, class Spec, class Grid, class DataStores>
                void run_fused_spec(Launcher &launcher, Spec, Grid const &grid, DataStores external) {
                    using msses_t = be_api::make_fused_view<Spec>;
                    static_assert(fusable<Spec>::value, "stencil::b200: spec cannot take the fused path");
                    // device memory behind the temporaries that are not (purely) cached
                    using tmp_plh_map_t = be_api::remove_caches_from_plh_map<
                        meta::filter<needs_memory, typename msses_t::tmp_plh_map_t>>;
                    auto alloc = launcher.allocator();
                    auto temporaries = be_api::make_data_stores(tmp_plh_map_t(), [&](auto info) {
                        using extent_t = decltype(info.extent());
                        return make_temporary<Geo, decltype(info.data())>(
                            has_ij_extent<extent_t>(), extent_t(), grid, msses_t::interval(), alloc);
                    });
                    auto fields = tuple_util::transform(
                        [](auto &&store) {
                            return sid::block(std::forward<decltype(store)>(store),
                                hymap::keys<dim::i, dim::j>::make_values(
                                    integral_constant<int_t, Geo::bi>(), integral_constant<int_t, Geo::bj>()));
                        },
                        std::move(external));
                    auto data_stores = hymap::concat(std::move(fields), std::move(temporaries));
                    using volatile_t = meta::if_c<Geo::chain_sweeps, written_in_sweeps<msses_t>, meta::list<>>;
                    launch_msses<Launcher, Geo, volatile_t>(
                        launcher, meta::rename<meta::list, msses_t>(), grid, data_stores, nothing_pending());
                }

#ifdef __CUDACC__
                // ---------------------------------------------------------------- the CUDA side
                struct cuda_cta {
                    static GT_FUNCTION int_t tid() {
#ifdef __CUDA_ARCH__
                        return threadIdx.x;
#else
                        return 0;
#endif
                    }
                    static GT_FUNCTION int_t block_i() {
#ifdef __CUDA_ARCH__
                        return blockIdx.x;
#else
                        return 0;
#endif
                    }
                    static GT_FUNCTION int_t block_j() {
#ifdef __CUDA_ARCH__
                        return blockIdx.y;
#else
                        return 0;
#endif
                    }
                    static GT_FUNCTION int_t block_k() {
#ifdef __CUDA_ARCH__
                        return blockIdx.z;
#else
                        return 0;
#endif
                    }
                    static GT_FUNCTION void sync() {
#ifdef __CUDA_ARCH__
                        __syncthreads();
#endif
                    }
                    static GT_FUNCTION char *smem() {
#ifdef __CUDA_ARCH__
                        extern __shared__ __align__(16) char gtb200_fused_smem[];
                        return gtb200_fused_smem;
#else
                        return nullptr;
#endif
                    }
                };

                template <class Body>
                __global__ void __launch_bounds__(Body::threads) mss_kernel(const __grid_constant__ Body body) {
                    body();
                }

                struct cuda_launcher {
                    using cta_t = cuda_cta;
                    cudaStream_t m_stream;

                    /// TMA descriptor of a 3-d box (libgtb200 encodes it); false if the field is not TMA-addressable
                    static bool tensor_map(void *map, const void *base, int elem_size, const int64_t dims[3],
                        const int64_t strides_bytes[2], const int box[3]) {
                        return gtb_tensor_map_3d(map, base, elem_size, dims, strides_bytes, box) == GTB_OK;
                    }

                    // Device memory for the temporaries, recycled like the reference's sid::device::cached_allocator
                    // (sid/allocator.hpp:65-95: thread-local free lists by size) -- but the free lists are per STREAM:
                    // a block is handed back as soon as run() returns, while the kernels that use it may still be in
                    // flight.  On the same stream the next user is ordered behind them; a launch on another stream
                    // (b200<stream_getter>) is not, so it must not get that block.
                    struct stream_cached_malloc {
                        cudaStream_t m_stream;
                        struct cache_t {
                            std::map<std::pair<cudaStream_t, size_t>, std::vector<char *>> free;
                            ~cache_t() {
                                for (auto &kv : free)
                                    for (char *p : kv.second)
                                        cudaFree(p);
                            }
                        };
                        static cache_t &cache() {
                            static thread_local cache_t c;
                            return c;
                        }
                        struct deleter_f {
                            cudaStream_t m_stream;
                            size_t m_size;
                            void operator()(char *p) const { cache().free[{m_stream, m_size}].push_back(p); }
                        };
                        std::unique_ptr<char[], deleter_f> operator()(size_t size) const {
                            auto &list = cache().free[{m_stream, size}];
                            char *p = nullptr;
                            if (list.empty()) {
                                GT_CUDA_CHECK(cudaMalloc(&p, size));
                            } else {
                                p = list.back();
                                list.pop_back();
                            }
                            return std::unique_ptr<char[], deleter_f>(p, deleter_f{m_stream, size});
                        }
                    };
                    auto allocator() const { return sid::device::allocator(stream_cached_malloc{m_stream}); }

                    template <class Body>
                    void launch(Body const &body, int_t nbi, int_t nbj, int_t nbk, int_t threads, int_t smem) const {
                        if (smem > 48 * 1024)
                            GT_CUDA_CHECK(cudaFuncSetAttribute(
                                mss_kernel<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                        mss_kernel<Body><<<dim3(nbi, nbj, nbk), dim3(threads), smem, m_stream>>>(body);
                        GT_CUDA_CHECK(cudaGetLastError());
                    }
                };
#endif
            } // namespace fused
        } // namespace b200_backend
    } // namespace stencil
} // namespace gridtools
