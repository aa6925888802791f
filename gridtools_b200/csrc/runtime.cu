// runtime.cu -- error state, options, device state, cached scratch and the TMA descriptor encoder of libgtb200.
#include "common.cuh"

#include <thread>

#include <cstring>

#include <map>

#include <mutex>
#include <string>

namespace gtb {

    namespace {
        thread_local char t_error[512] = "";
        std::mutex g_mutex;
        constexpr int max_devices = 64;
        device_state g_dev[max_devices];
        struct slab {
            void *ptr = nullptr;
            size_t bytes = 0;
        };
        // one slab per (device, stream): launches on one stream run in order and may share a slab, launches on
        // different streams (or from different host threads on different streams) may overlap and must not
        std::map<std::pair<int, cudaStream_t>, slab> g_scratch;
        options g_opts;
    } // namespace

    std::atomic<int64_t> g_launches{0};

    void set_error(const char *fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(t_error, sizeof(t_error), fmt, ap);
        va_end(ap);
    }

    int fail(gtb_status st, const char *fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(t_error, sizeof(t_error), fmt, ap);
        va_end(ap);
        return st;
    }

    int cuda_fail(cudaError_t e, const char *what) {
        snprintf(t_error, sizeof(t_error), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
        return GTB_ERR_CUDA;
    }

    options &opts() { return g_opts; }

    device_state *dev() {
        int d = -1;
        cudaError_t e = cudaGetDevice(&d);
        if (e != cudaSuccess || d < 0 || d >= max_devices) {
            cuda_fail(e, "cudaGetDevice (libgtb200 has no CPU fallback: a CUDA device is required)");
            return nullptr;
        }
        device_state &s = g_dev[d];
        if (s.device == d)
            return &s;
        std::lock_guard<std::mutex> lock(g_mutex);
        if (s.device == d)
            return &s;
        cudaDeviceProp p;
        e = cudaGetDeviceProperties(&p, d);
        if (e != cudaSuccess) {
            cuda_fail(e, "cudaGetDeviceProperties");
            return nullptr;
        }
        if (p.major != 10) {
            set_error("libgtb200 is built for sm_100a only; device %d is sm_%d%d", d, p.major, p.minor);
            return nullptr;
        }
        s.sm_count = p.multiProcessorCount;
        s.l2_bytes = p.l2CacheSize;
        s.hbm_bytes = (int64_t)p.totalGlobalMem;
        s.max_smem_optin = (int)p.sharedMemPerBlockOptin;
        s.persisting_l2_max = (int64_t)p.persistingL2CacheMaxSize;
        s.device = d;
        return &s;
    }

    void *scratch(size_t bytes, cudaStream_t stream) {
        device_state *s = dev();
        if (!s)
            return nullptr;
        std::lock_guard<std::mutex> lock(g_mutex);
        slab &sl = g_scratch[{s->device, stream}];
        if (sl.bytes >= bytes && sl.ptr)
            return sl.ptr;
        if (sl.ptr) {
            cudaDeviceSynchronize(); // earlier kernels may still use the old slab
            cudaFree(sl.ptr);
            sl = slab{};
        }
        size_t want = bytes + bytes / 4; // head room so slightly larger domains do not reallocate
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            cuda_fail(e, "cudaMalloc(scratch)");
            return nullptr;
        }
        sl.ptr = p;
        sl.bytes = want;
        return p;
    }

    encode_tiled_fn tensor_map_encoder() {
        static encode_tiled_fn fn = [] {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult qres;
            cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
            if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess)
                p = nullptr;
            return reinterpret_cast<encode_tiled_fn>(p);
        }();
        return fn;
    }

} // namespace gtb

using namespace gtb;

GTB_API int gtb_version(void) { return GTB_VERSION; }

GTB_API const char *gtb_last_error(void) { return t_error; }

GTB_API int gtb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

GTB_API int gtb_init(int device) {
    GTB_CUDA(cudaSetDevice(device));
    if (!dev())
        return GTB_ERR_CUDA;
    return GTB_OK;
}

GTB_API int gtb_device_info(int *sm_count, int64_t *l2_bytes, int64_t *hbm_bytes) {
    device_state *s = dev();
    if (!s)
        return GTB_ERR_CUDA;
    if (sm_count)
        *sm_count = s->sm_count;
    if (l2_bytes)
        *l2_bytes = s->l2_bytes;
    if (hbm_bytes)
        *hbm_bytes = s->hbm_bytes;
    return GTB_OK;
}

namespace {
    int *find_option(const char *key) {
        if (!key)
            return nullptr;
        options &o = opts();
        struct {
            const char *name;
            int *ptr;
        } table[] = {{"hd.variant", &o.hd_variant},
            {"hd.stages", &o.hd_stages},
            {"hd.ctas_per_sm", &o.hd_ctas_per_sm},
            {"va.variant", &o.va_variant},
            {"va.threads", &o.va_threads},
            {"va.unroll", &o.va_unroll},
            {"va.scratch", &o.va_scratch},
            {"va.hints", &o.va_hints},
            {"va.ctas_per_sm", &o.va_ctas_per_sm},
            {"va.save_upos", &o.va_save_upos},
            {"l2.persist_mb", &o.l2_persist_mb},
            {"va.debug", &o.va_debug},
            {"va.stages", &o.va_stages},
            {"va.stagger", &o.va_stagger},
            {"va.bldg", &o.va_bldg},
            {"reserve_sms", &o.reserve_sms},
            {"pdl", &o.pdl},
            {"halo.fused", &o.halo_fused},
            {"halo.dma", &o.halo_dma},
            {"halo.timeout_ms", &o.halo_timeout_ms},
            {"halo.vec", &o.halo_vec},
            {"halo.max_blocks", &o.halo_max_blocks},
            {"copy.vec", &o.copy_vec}};
        for (auto &t : table)
            if (strcmp(t.name, key) == 0)
                return t.ptr;
        return nullptr;
    }
} // namespace

GTB_API int gtb_set_option(const char *key, int value) {
    int *p = find_option(key);
    if (!p)
        return fail(GTB_ERR_ARG, "unknown option '%s'", key ? key : "(null)");
    *p = value;
    return GTB_OK;
}

GTB_API int gtb_get_option(const char *key, int *value) {
    int *p = find_option(key);
    if (!p || !value)
        return fail(GTB_ERR_ARG, "unknown option '%s'", key ? key : "(null)");
    *value = *p;
    return GTB_OK;
}

namespace gtb {
    // L2 set-aside for persisting (evict_last) accesses: the k-cache slabs of the vertical sweeps live there.
    int set_l2_persist(int64_t bytes) {
        device_state *s = dev();
        if (!s)
            return GTB_ERR_CUDA;
        if (bytes > s->persisting_l2_max)
            bytes = s->persisting_l2_max;
        if (bytes < 0)
            bytes = 0;
        if (s->persisting_l2_set == bytes)
            return GTB_OK;
        GTB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)bytes));
        s->persisting_l2_set = bytes;
        return GTB_OK;
    }
} // namespace gtb

GTB_API int gtb_release_scratch(void) {
    device_state *s = dev();
    if (!s)
        return GTB_ERR_CUDA;
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaDeviceSynchronize();
    for (auto it = g_scratch.begin(); it != g_scratch.end();) {
        if (it->first.first == s->device) {
            cudaFree(it->second.ptr);
            it = g_scratch.erase(it);
        } else
            ++it;
    }
    return GTB_OK;
}

namespace gtb {
    const char *&last_kernel_name() {
        static thread_local const char *name = "";
        return name;
    }
} // namespace gtb

GTB_API const char *gtb_last_kernel(void) { return gtb::last_kernel_name(); }

GTB_API int64_t gtb_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------ streams for hosts
// Host bindings that do not want to include the CUDA runtime (the gcl handler of include/gtb200/gcl/b200.hpp) get
// their private stream here.  gtb_stream_create gives a NON-BLOCKING stream: work on it is not ordered against the
// legacy default stream implicitly, gtb_stream_after_default() orders it explicitly after everything issued to the
// legacy default stream so far (what a kernel launched on a blocking stream would get for free).
GTB_API int gtb_stream_create(void **stream, int high_priority) {
    if (!stream)
        return fail(GTB_ERR_ARG, "gtb_stream_create: null argument");
    if (!dev())
        return GTB_ERR_CUDA;
    int lo = 0, hi = 0;
    GTB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s;
    GTB_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
    *stream = s;
    return GTB_OK;
}

GTB_API int gtb_stream_destroy(void *stream) {
    if (stream)
        GTB_CUDA(cudaStreamDestroy(as_stream(stream)));
    return GTB_OK;
}

GTB_API int gtb_stream_after_default(void *stream) {
    cudaEvent_t e;
    GTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaError_t st = cudaEventRecord(e, cudaStreamLegacy);
    if (st == cudaSuccess)
        st = cudaStreamWaitEvent(as_stream(stream), e, 0);
    cudaEventDestroy(e); // released when the record has completed
    if (st != cudaSuccess)
        return cuda_fail(st, "gtb_stream_after_default");
    return GTB_OK;
}

GTB_API int gtb_stream_synchronize(void *stream) {
    GTB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return GTB_OK;
}

// ------------------------------------------------------------------------------------------------ tensor maps
// For kernels that are compiled in the USER's translation unit (the generic fused path of stencil::b200 instantiates
// the user's functors there): a TMA descriptor of a 3-d box over an i-contiguous field, encoded here so that the
// header needs neither libcuda nor the driver entry point query.  GTB_ERR_LAYOUT when the field is not
// TMA-addressable (base or strides not multiples of 16 bytes, box row not a multiple of 16 bytes, box too large).
GTB_API int gtb_tensor_map_3d(void *map128, const void *base, int elem_size, const int64_t dims[3],
    const int64_t strides_bytes[2], const int box[3]) {
    if (!map128 || !base || !dims || !strides_bytes || !box)
        return fail(GTB_ERR_ARG, "gtb_tensor_map_3d: null argument");
    if (!dev())
        return GTB_ERR_CUDA;
    auto enc = tensor_map_encoder();
    if (!enc)
        return fail(GTB_ERR_CUDA, "gtb_tensor_map_3d: cuTensorMapEncodeTiled not available");
    CUtensorMapDataType dt;
    if (elem_size == 8)
        dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    else if (elem_size == 4)
        dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    else
        return fail(GTB_ERR_LAYOUT, "gtb_tensor_map_3d: element size %d", elem_size);
    if (reinterpret_cast<uintptr_t>(base) % 16 || strides_bytes[0] % 16 || strides_bytes[1] % 16 ||
        strides_bytes[0] <= 0 || strides_bytes[1] <= 0 || (box[0] * (int64_t)elem_size) % 16 || box[0] > 256 ||
        box[1] > 256 || box[2] > 256 || box[0] < 1 || box[1] < 1 || box[2] < 1 || dims[0] < 1 || dims[1] < 1 ||
        dims[2] < 1)
        return fail(GTB_ERR_LAYOUT, "gtb_tensor_map_3d: field is not TMA-addressable");
    cuuint64_t d[3] = {(cuuint64_t)dims[0], (cuuint64_t)dims[1], (cuuint64_t)dims[2]};
    cuuint64_t st[2] = {(cuuint64_t)strides_bytes[0], (cuuint64_t)strides_bytes[1]};
    cuuint32_t bx[3] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2]};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(static_cast<CUtensorMap *>(map128), dt, 3, const_cast<void *>(base), d, st, bx, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(GTB_ERR_LAYOUT, "gtb_tensor_map_3d: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return GTB_OK;
}

// ------------------------------------------------------------------------------------------------ staged copies
// Whole-allocation transfers between PAGEABLE host memory and the device, for storage traits (include/gtb200/storage/
// b200.hpp): the host mirror of a GridTools data_store is a plain new[] array that the traits never see allocated or
// freed (storage/data_store.hpp:101-104), so it cannot be pinned in place.  A plain cudaMemcpy from pageable memory
// (storage/gpu.hpp:86-99) lets the driver stage through its own small pinned buffers at the speed of ONE host thread;
// here the chunks are staged by several host threads into a ring of pinned buffers and sent by asynchronous copies that
// overlap the staging of the next chunk.  Stream-ordered on `stream` (nullptr = legacy default stream, like the reference).
namespace {
    constexpr size_t kStageChunk = 8u << 20;
    constexpr int kStageSlots = 3, kStageThreads = 4;
    struct stage_ring {
        char *slot[kStageSlots] = {};
        cudaEvent_t done[kStageSlots] = {};
        int device = -1;
        ~stage_ring() {
            for (int i = 0; i < kStageSlots; ++i)
                if (slot[i]) {
                    cudaFreeHost(slot[i]);
                    cudaEventDestroy(done[i]);
                }
        }
    };
    stage_ring *ring() {
        static thread_local stage_ring r;
        device_state *d = dev();
        if (!d)
            return nullptr;
        if (!r.slot[0]) {
            for (int i = 0; i < kStageSlots; ++i) {
                if (cudaHostAlloc(&r.slot[i], kStageChunk, cudaHostAllocDefault) != cudaSuccess ||
                    cudaEventCreateWithFlags(&r.done[i], cudaEventDisableTiming) != cudaSuccess) {
                    cuda_fail(cudaGetLastError(), "staged copy: pinned ring");
                    return nullptr;
                }
            }
            r.device = d->device;
        }
        return &r;
    }
    void parallel_copy(char *dst, const char *src, size_t n) {
        if (n < (1u << 20)) {
            std::memcpy(dst, src, n);
            return;
        }
        std::thread th[kStageThreads - 1];
        const size_t part = (n / kStageThreads + 63) & ~size_t(63);
        for (int t = 0; t < kStageThreads - 1; ++t) {
            const size_t off = (size_t)(t + 1) * part;
            if (off < n)
                th[t] = std::thread([=] { std::memcpy(dst + off, src + off, off + part < n ? part : n - off); });
        }
        std::memcpy(dst, src, part < n ? part : n);
        for (auto &t : th)
            if (t.joinable())
                t.join();
    }
} // namespace

namespace {
    // page-locked host memory (gtb_host_malloc, cudaHostAlloc, cudaHostRegister)?
    bool is_pinned_host(const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    }
} // namespace

GTB_API int gtb_host_malloc(void **out, int64_t bytes) {
    if (!out || bytes < 0)
        return fail(GTB_ERR_ARG, "gtb_host_malloc: bad argument");
    if (!dev())
        return GTB_ERR_CUDA;
    *out = nullptr;
    if (bytes == 0)
        return GTB_OK;
    cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cuda_fail(e, "gtb_host_malloc");
        return GTB_ERR_ALLOC;
    }
    return GTB_OK;
}

GTB_API int gtb_host_free(void *p) {
    if (p)
        cudaFreeHost(p);
    return GTB_OK;
}

GTB_API int gtb_device_malloc(void **out, int64_t bytes) {
    if (!out || bytes < 0)
        return fail(GTB_ERR_ARG, "gtb_device_malloc: bad argument");
    if (!dev())
        return GTB_ERR_CUDA;
    *out = nullptr;
    if (bytes == 0)
        return GTB_OK;
    cudaError_t e = cudaMalloc(out, (size_t)bytes);
    if (e != cudaSuccess) {
        cuda_fail(e, "gtb_device_malloc");
        return GTB_ERR_ALLOC;
    }
    return GTB_OK;
}

GTB_API int gtb_device_free(void *p) {
    if (p)
        cudaFree(p);
    return GTB_OK;
}

GTB_API int gtb_staged_upload(void *device_dst, const void *host_src, int64_t bytes, void *stream) {
    if (bytes < 0 || (bytes && (!device_dst || !host_src)))
        return fail(GTB_ERR_ARG, "gtb_staged_upload: bad argument");
    if (!dev())
        return GTB_ERR_CUDA;
    cudaStream_t s = stream ? as_stream(stream) : cudaStreamLegacy;
    if (bytes && is_pinned_host(host_src)) { // a pinned mirror (gtb_host_malloc): the copy engine reads it directly
        GTB_CUDA(cudaMemcpyAsync(device_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, s));
        GTB_CUDA(cudaStreamSynchronize(s)); // host_src may be written again when this returns
        return GTB_OK;
    }
    stage_ring *r = ring();
    if (!r)
        return GTB_ERR_CUDA;
    int i = 0;
    for (int64_t off = 0; off < bytes; off += (int64_t)kStageChunk, i = (i + 1) % kStageSlots) {
        const size_t n = (size_t)(bytes - off < (int64_t)kStageChunk ? bytes - off : (int64_t)kStageChunk);
        GTB_CUDA(cudaEventSynchronize(r->done[i])); // the copy that last used this slot has drained it
        parallel_copy(r->slot[i], static_cast<const char *>(host_src) + off, n);
        GTB_CUDA(cudaMemcpyAsync(static_cast<char *>(device_dst) + off, r->slot[i], n, cudaMemcpyHostToDevice, s));
        GTB_CUDA(cudaEventRecord(r->done[i], s));
    }
    return GTB_OK; // host_src may be modified again; the device copy is complete in stream order
}

GTB_API int gtb_staged_download(void *host_dst, const void *device_src, int64_t bytes, void *stream) {
    if (bytes < 0 || (bytes && (!host_dst || !device_src)))
        return fail(GTB_ERR_ARG, "gtb_staged_download: bad argument");
    if (!dev())
        return GTB_ERR_CUDA;
    cudaStream_t s = stream ? as_stream(stream) : cudaStreamLegacy;
    if (bytes && is_pinned_host(host_dst)) {
        GTB_CUDA(cudaMemcpyAsync(host_dst, device_src, (size_t)bytes, cudaMemcpyDeviceToHost, s));
        GTB_CUDA(cudaStreamSynchronize(s));
        return GTB_OK;
    }
    stage_ring *r = ring();
    if (!r)
        return GTB_ERR_CUDA;
    const int64_t n_chunks = (bytes + (int64_t)kStageChunk - 1) / (int64_t)kStageChunk;
    auto size_of = [&](int64_t c) {
        const int64_t off = c * (int64_t)kStageChunk;
        return (size_t)(bytes - off < (int64_t)kStageChunk ? bytes - off : (int64_t)kStageChunk);
    };
    auto issue = [&](int64_t c) -> int {
        const int i = (int)(c % kStageSlots);
        GTB_CUDA(cudaMemcpyAsync(r->slot[i], static_cast<const char *>(device_src) + c * (int64_t)kStageChunk, size_of(c),
            cudaMemcpyDeviceToHost, s));
        GTB_CUDA(cudaEventRecord(r->done[i], s));
        return GTB_OK;
    };
    for (int64_t c = 0; c < n_chunks && c < kStageSlots - 1; ++c) // prime the ring
        if (int st = issue(c))
            return st;
    for (int64_t c = 0; c < n_chunks; ++c) {
        if (c + kStageSlots - 1 < n_chunks)
            if (int st = issue(c + kStageSlots - 1))
                return st;
        const int i = (int)(c % kStageSlots);
        GTB_CUDA(cudaEventSynchronize(r->done[i]));
        parallel_copy(static_cast<char *>(host_dst) + c * (int64_t)kStageChunk, r->slot[i], size_of(c));
    }
    return GTB_OK; // host_dst is complete
}

// ------------------------------------------------------------------------------------------------ stencil gates
namespace gtb {
    namespace {
        thread_local stencil_gate t_gate;
        std::atomic<unsigned> g_gate_launch{0};
    } // namespace

    int *gate_counter_slot() {
        static int *base[64] = {};
        static std::mutex mtx;
        std::lock_guard<std::mutex> lock(mtx);
        device_state *d = dev();
        if (!d || d->device < 0 || d->device >= 64)
            return nullptr;
        if (!base[d->device]) {
            int *q = nullptr;
            if (cudaMalloc(&q, 64 * sizeof(int)) != cudaSuccess || cudaMemset(q, 0, 64 * sizeof(int)) != cudaSuccess) {
                cuda_fail(cudaGetLastError(), "gate counters");
                return nullptr;
            }
            base[d->device] = q;
        }
        return base[d->device] + (g_gate_launch.fetch_add(1, std::memory_order_relaxed) % 16);
    }

    unsigned long long *gate_timeout_counter() {
        static unsigned long long *ctr[64] = {};
        static std::mutex mtx;
        std::lock_guard<std::mutex> lock(mtx);
        device_state *d = dev();
        if (!d || d->device < 0 || d->device >= 64)
            return nullptr;
        if (!ctr[d->device]) {
            unsigned long long *q = nullptr;
            if (cudaMalloc(&q, sizeof(*q)) != cudaSuccess || cudaMemset(q, 0, sizeof(*q)) != cudaSuccess) {
                cuda_fail(cudaGetLastError(), "gate timeout counter");
                return nullptr;
            }
            ctr[d->device] = q;
        }
        return ctr[d->device];
    }

    stencil_gate take_gate() {
        stencil_gate g = t_gate;
        t_gate = stencil_gate();
        if (g.post && !g.cta_done)
            g.cta_done = gate_counter_slot();
        if (g.wait_flag)
            g.timeouts = gate_timeout_counter();
        return g;
    }
} // namespace gtb

GTB_API int gtb_stencil_gate(const void *wait_flag, uint64_t wait_value, void *post_counter) {
    if (!gtb::dev())
        return GTB_ERR_CUDA;
    if ((wait_flag || post_counter) && gtb::opts().reserve_sms < 1)
        return gtb::fail(GTB_ERR_STATE,
            "gtb_stencil_gate: a gated stencil spins on the device until another kernel raises its flag; that kernel needs "
            "SMs the stencil does not occupy -- set option reserve_sms >= 1 first");
    gtb::t_gate.wait_flag = static_cast<const unsigned long long *>(wait_flag);
    gtb::t_gate.wait_value = wait_value;
    gtb::t_gate.post = static_cast<unsigned long long *>(post_counter);
    gtb::t_gate.cta_done = nullptr;
    return GTB_OK;
}

GTB_API int gtb_gate_timeouts(int64_t *count) {
    if (!count)
        return gtb::fail(GTB_ERR_ARG, "gtb_gate_timeouts: null argument");
    if (!gtb::dev())
        return GTB_ERR_CUDA;
    unsigned long long *c = gtb::gate_timeout_counter();
    if (!c)
        return GTB_ERR_ALLOC;
    unsigned long long v = 0;
    GTB_CUDA(cudaDeviceSynchronize());
    GTB_CUDA(cudaMemcpy(&v, c, sizeof(v), cudaMemcpyDeviceToHost));
    *count = (int64_t)v;
    return GTB_OK;
}

// ------------------------------------------------------------------------------------------------ box copies
// Host <-> device copy of a sub-box of a 3-d field (i contiguous) on a stream: what storage::gpu's update_target /
// update_host (storage/gpu.hpp:86-99: one blocking cudaMemcpy of the whole padded allocation) would move if it only
// moved the points a stencil reads or writes.  `host` must be pinned for the copy to be asynchronous.
GTB_API int gtb_copy_box_async(void *device_origin, void *host_origin, int elem_size, int64_t stride_j, int64_t stride_k,
    int ni, int nj, int nk, int to_device, void *stream) {
    if (!device_origin || !host_origin || (elem_size != 4 && elem_size != 8) || ni < 0 || nj < 0 || nk < 0 ||
        stride_j < ni || stride_k < stride_j * nj)
        return gtb::fail(GTB_ERR_ARG, "gtb_copy_box_async: bad argument");
    if (!gtb::dev())
        return GTB_ERR_CUDA;
    if (ni == 0 || nj == 0 || nk == 0)
        return GTB_OK;
    if (stride_k % stride_j != 0) // cudaMemcpy3D needs whole rows per slice: fall back to one 2-d copy per level
    {
        for (int k = 0; k < nk; ++k) {
            char *d = static_cast<char *>(device_origin) + (size_t)k * stride_k * elem_size;
            char *h = static_cast<char *>(host_origin) + (size_t)k * stride_k * elem_size;
            GTB_CUDA(cudaMemcpy2DAsync(to_device ? d : h, (size_t)stride_j * elem_size, to_device ? h : d,
                (size_t)stride_j * elem_size, (size_t)ni * elem_size, (size_t)nj,
                to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, gtb::as_stream(stream)));
        }
        return GTB_OK;
    }
    cudaMemcpy3DParms prm = {};
    const size_t pitch = (size_t)stride_j * elem_size, rows = (size_t)(stride_k / stride_j);
    cudaPitchedPtr dp = make_cudaPitchedPtr(device_origin, pitch, pitch, rows);
    cudaPitchedPtr hp = make_cudaPitchedPtr(host_origin, pitch, pitch, rows);
    prm.srcPtr = to_device ? hp : dp;
    prm.dstPtr = to_device ? dp : hp;
    prm.extent = make_cudaExtent((size_t)ni * elem_size, (size_t)nj, (size_t)nk);
    prm.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    GTB_CUDA(cudaMemcpy3DAsync(&prm, gtb::as_stream(stream)));
    return GTB_OK;
}
