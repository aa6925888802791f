// common.cuh -- shared host/device helpers of libgtb200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <utility>

#include "../../include/gtb200.h"

#define GTB_API extern "C" __attribute__((visibility("default")))

namespace gtb {

    // ---------------------------------------------------------------- errors (thread local, never throws)
    void set_error(const char *fmt, ...);
    int fail(gtb_status st, const char *fmt, ...);
    int cuda_fail(cudaError_t e, const char *what);

#define GTB_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess)                          \
            return ::gtb::cuda_fail(e__, #call);         \
    } while (0)

    // ---------------------------------------------------------------- options
    struct options {
        int hd_variant = 0;   // 0 auto, 1 cp.async staged, 2 TMA + block barrier
        int hd_stages = 0;    // 0 auto
        int hd_ctas_per_sm = 0;
        int va_variant = 0;   // 0 auto, 1 first TMA version, 3 TMA + L2 slab (fp32 default), 7 paired warps + tensor memory (fp64 default)
        int va_threads = 0;   // threads per CTA (multiple of 32)
        int va_unroll = 0;    // k levels prefetched ahead
        int va_scratch = 0;   // 0 auto, 1 global (L2) scratch, 2 shared memory
        int va_hints = 1;     // L2 eviction-priority hints on/off
        int va_ctas_per_sm = 0; // > 0: persistent grid of that many CTAs per SM with per-thread scratch slots
        int va_save_upos = 0; // u_pos(k) kept next to ccol/dcol instead of re-read in the backward sweep: 0 auto, 1 on, 2 off
        int va_debug = 0;     // diagnosis only, see va_params::debug
        int va_stages = 0;    // TMA ring depth (0 auto)
        int va_bldg = 0;      // paired-warp kernel: u_pos of the backward sweep 0 auto, 1 register loads (LDG), 2 TMA ring
        int va_stagger = 0;   // TMEM variant: start stagger between the warps of a CTA, in units of 100 ns
        int copy_vec = 1;     // vectorised copy on/off
        int l2_persist_mb = -1; // L2 set-aside for the k-cache slabs in MB: -1 auto (slab size), 0 off
        int halo_max_blocks = 0; // > 0: grid size cap of the halo transfer kernels (0: one block per SM)
        int halo_vec = 1;       // 16-byte vector transfers where the halo regions allow it
        int halo_timeout_ms = 60000; // device-side waits for a neighbour's message give up after this long; 0: never
        int halo_dma = 0;       // gtb_halo_exchange / gtb_halo_send: the NVLink leg as copy-engine transfers of the packed messages
        int halo_fused = 0;     // gtb_halo_exchange as ONE launch (pack, signal, wait, unpack); 0: two launches
        int pdl = 1;            // programmatic dependent launch of the vertical advection kernel (prologue under the previous kernel's tail): 0 off, 1 unless SMs are reserved, 2 always
        int reserve_sms = 0;    // SMs the persistent stencil grids leave free (for a halo exchange that runs beside them)
    };
    options &opts();

    // ---------------------------------------------------------------- device info / scratch
    struct device_state {
        int device = -1;
        int sm_count = 0;
        int64_t l2_bytes = 0;
        int64_t hbm_bytes = 0;
        int max_smem_optin = 0;
        int64_t persisting_l2_max = 0;
        int64_t persisting_l2_set = -1;
    };
    // Lazily initialised state of the current device; nullptr + error set if there is no usable device.
    device_state *dev();

    // Cached scratch (temporaries, the counterpart of the reference's sid::device::cached_allocator): one growing slab
    // per (device, stream).  Launches on one stream are ordered and share their slab; launches on different streams
    // may overlap, so every stream has its own (gtb_release_scratch frees them all).
    void *scratch(size_t bytes, cudaStream_t stream);
    int set_l2_persist(int64_t bytes);

    // Device-side ordering of a stencil launch against a halo exchange that runs beside it on another stream
    // (gtb_stencil_gate, include/gtb200.h): the kernel waits until *wait_flag >= wait_value before it touches global
    // memory and its last CTA adds 1 to *post when every CTA is done.  `cta_done` is a zeroed, self-resetting counter.
    struct stencil_gate {
        const unsigned long long *wait_flag = nullptr;
        unsigned long long wait_value = 0;
        unsigned long long *post = nullptr;
        int *cta_done = nullptr;
        unsigned long long *timeouts = nullptr; // device counter of waits that gave up (gtb_gate_timeouts)
    };
    unsigned long long *gate_timeout_counter(); // per device, zeroed once
    // The gate armed for the next gated-capable stencil launch of this thread (consumed by it); empty otherwise.
    stencil_gate take_gate();
    int *gate_counter_slot(); // one of 16 rotating zeroed ints per device

    extern std::atomic<int64_t> g_launches;
    inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

    inline cudaStream_t as_stream(void *s) { return static_cast<cudaStream_t>(s); }

    const char *&last_kernel_name(); // per thread: what the last check_launch was called with (gtb_last_kernel)
    inline int check_launch(const char *what) {
        last_kernel_name() = what;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return cuda_fail(e, what);
        return GTB_OK;
    }

    inline bool field_ok(const gtb_field *f) { return f && f->ptr; }

    // Launch with the programmatic-stream-serialization attribute (option "pdl"): the kernel must call ptx::pdl_wait()
    // before it reads or writes anything an earlier kernel of the stream may have touched.
    template <class... KArgs, class... Args>
    inline cudaError_t launch_pdl_if(bool allowed, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
        cudaStream_t stream, Args &&...args) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = allowed ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    }
    // option "pdl": 0 off, 1 unless SMs are reserved for an exchange on another stream (the early CTAs of the NEXT
    // launch would settle exactly on the reserved SMs), 2 always
    inline bool pdl_allowed() { return opts().pdl == 2 || (opts().pdl == 1 && opts().reserve_sms == 0); }
    template <class... KArgs, class... Args>
    inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
        Args &&...args) {
        return launch_pdl_if(pdl_allowed(), kernel, grid, block, smem, stream, std::forward<Args>(args)...);
    }

    // TMA descriptor encoder, resolved from the driver at run time (no link-time dependency on libcuda).
    typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    encode_tiled_fn tensor_map_encoder();

    constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

    // SMs a persistent stencil grid is sized for: all of them, less the ones reserved for a concurrent halo exchange.
    inline int stencil_sms(const device_state *d) {
        const int r = opts().reserve_sms;
        return r > 0 && r < d->sm_count ? d->sm_count - r : d->sm_count;
    }

} // namespace gtb

// ---------------------------------------------------------------------- device-side PTX wrappers
namespace gtb {
    namespace ptx {
#ifdef __CUDACC__
        __device__ __forceinline__ unsigned long long globaltimer() {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            return t;
        }
        __device__ __forceinline__ uint32_t smem_addr(const void *p) {
            return static_cast<uint32_t>(__cvta_generic_to_shared(p));
        }
        __device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
        }
        __device__ __forceinline__ void fence_barrier_init() {
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
        __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                         : "memory");
        }
        __device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
        }
        __device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
            uint32_t ok;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(smem_addr(bar)), "r"(parity)
                : "memory");
            return ok != 0;
        }
        __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
            while (!mbar_try_wait(bar, parity)) {
            }
        }
        // 3-D tiled TMA load global -> shared, completion on an mbarrier.
        __device__ __forceinline__ void tma_load_3d(
            void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_addr(dst)),
                "l"(reinterpret_cast<uint64_t>(map)),
                "r"(smem_addr(bar)),
                "r"(c0),
                "r"(c1),
                "r"(c2)
                : "memory");
        }
        __device__ __forceinline__ void tma_load_3d_hint(
            void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, uint64_t policy) {
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_addr(dst)),
                "l"(reinterpret_cast<uint64_t>(map)),
                "r"(smem_addr(bar)),
                "r"(c0),
                "r"(c1),
                "r"(c2),
                "l"(policy)
                : "memory");
        }
        // True in exactly one (the lowest active) lane of a converged warp.  Guarding TMA issue with this instead of
        // `lane == 0` tells ptxas that a single thread executes the region, so UTMALDG's uniform operands need no
        // per-lane election loop.
        __device__ __forceinline__ bool elect_one() {
            uint32_t pred;
            asm volatile(
                "{\n\t.reg .pred P;\n\t"
                "elect.sync _|P, 0xffffffff;\n\t"
                "selp.u32 %0, 1, 0, P;\n\t}"
                : "=r"(pred));
            return pred != 0;
        }
        // 1-D bulk copy global -> shared (contiguous, 16-byte aligned address and size), completion on an mbarrier.
        __device__ __forceinline__ void bulk_load_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
            uint64_t policy) {
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
                    "r"(smem_addr(dst)),
                "l"(src),
                "r"(bytes),
                "r"(smem_addr(bar)),
                "l"(policy)
                : "memory");
        }
        // Orders this thread's earlier generic-proxy accesses (all state spaces) before later async-proxy accesses
        // (TMA / bulk copies): needed when data written with ordinary stores is re-read by a bulk copy.
        __device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
        __device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
        }
        __device__ __forceinline__ uint64_t policy_evict_first() {
            uint64_t p;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
            return p;
        }
        __device__ __forceinline__ uint64_t policy_evict_last() {
            uint64_t p;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
            return p;
        }
        // cp.async (LDGSTS) with zero fill when !pred.
        template <int Bytes>
        __device__ __forceinline__ void cp_async(void *dst, const void *src, bool pred) {
            static_assert(Bytes == 4 || Bytes == 8 || Bytes == 16, "cp.async size");
            int n = pred ? Bytes : 0;
            if constexpr (Bytes == 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(n)
                             : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(smem_addr(dst)),
                             "l"(src),
                             "n"(Bytes),
                             "r"(n)
                             : "memory");
        }
        __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
        template <int N>
        __device__ __forceinline__ void cp_async_wait() {
            asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
        }

        // Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may be
        // scheduled while the previous kernel of the stream is still draining; pdl_wait() blocks until that kernel has
        // completed and its memory operations are visible (nothing it wrote may be touched before), pdl_launch_dependents()
        // lets the NEXT kernel of the stream start being scheduled as SMs free up.  Both are no-ops for a plain launch.
        __device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
        __device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

        __device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
            unsigned long long v;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
            return v;
        }
        // Spins (one thread) until *flag >= value; gives up after ~0.2 s so that a lost producer cannot hang the device,
        // and counts that in *timeouts (the host asks with gtb_gate_timeouts).
        __device__ __forceinline__ void gate_wait(const unsigned long long *flag, unsigned long long value,
            unsigned long long *timeouts) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(flag) < value) {
                if (clock64() - t0 > 400000000ll) {
                    if (timeouts)
                        atomicAdd(timeouts, 1ULL);
                    break;
                }
                __nanosleep(64);
            }
        }

        // Streaming global accesses with an L2 eviction policy.
        template <class T>
        __device__ __forceinline__ T ld_hint(const T *p, uint64_t policy) {
            T v;
            if constexpr (sizeof(T) == 8) {
                uint64_t r;
                asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(policy));
                memcpy(&v, &r, 8);
            } else {
                uint32_t r;
                asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
                memcpy(&v, &r, 4);
            }
            return v;
        }
        // Same, bypassing L1 (ld.global.cg): for data another warp of the SM has just rewritten through L2.
        template <class T>
        __device__ __forceinline__ T ld_cg_hint(const T *p, uint64_t policy) {
            T v;
            if constexpr (sizeof(T) == 8) {
                uint64_t r;
                asm volatile("ld.global.cg.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(policy) : "memory");
                memcpy(&v, &r, 8);
            } else {
                uint32_t r;
                asm volatile("ld.global.cg.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy) : "memory");
                memcpy(&v, &r, 4);
            }
            return v;
        }
        template <class T>
        __device__ __forceinline__ void st_hint(T *p, T v, uint64_t policy) {
            if constexpr (sizeof(T) == 8) {
                uint64_t r;
                memcpy(&r, &v, 8);
                asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(r), "l"(policy) : "memory");
            } else {
                uint32_t r;
                memcpy(&r, &v, 4);
                asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(r), "l"(policy) : "memory");
            }
        }
#endif
    } // namespace ptx
} // namespace gtb
