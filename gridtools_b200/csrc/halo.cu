// halo.cu -- gcl halo exchange for ranks on one NVLink/NVSwitch box (one process per GPU, or several "ranks" in one
// process for tests).
//
// Replaces:
//   * hndlr_dynamic_ut<..., gpu>::setup / pack / unpack        gcl/high_level/descriptors_manual_gpu.hpp:164-514
//   * the 24 per-direction, per-field kernels                   gcl/high_level/m_{pack,unpack}{X,Y,Z}{L,U}.hpp
//   * Halo_Exchange_3D's Irecv / Isend / Wait choreography      gcl/low_level/Halo_Exchange_3D.hpp:546-931
//
// Design:
//   * ONE launch packs every field for every neighbour (grid.y = direction, grid.z = field) instead of up to
//     6 x n_fields launches followed by cudaDeviceSynchronize (descriptors_manual_gpu.hpp:438-487);
//   * with the peer-to-peer transport the pack kernel stores straight into the neighbour's receive buffer through
//     the NVLink-mapped (cudaIpc) address -- there is no send buffer, no MPI and no host synchronisation; a
//     release-store of the epoch number into the neighbour's flag word publishes the message, the receiver's
//     wait kernel acquires it and the unpack kernel follows in stream order;
//   * receive buffers are double buffered by epoch parity: seeing epoch e+1 from a neighbour proves it finished
//     unpacking epoch e, so epoch e+2 may overwrite that buffer without any acknowledgement traffic;
//   * staged mode (send/recv buffers exposed as raw device pointers) lets a host move the messages with NCCL or MPI
//     instead.
// Index ranges follow common/halo_descriptor.hpp:90-201 (loop_{low,high}_bound_{inside,outside}, s_length,
// r_length); message layout is field-major, dimension 0 fastest (gcl/high_level/descriptors.hpp:61-91).
#include "common.cuh"

#include <unistd.h>

using namespace gtb;

namespace {

    constexpr int kMaxFields = 16; // per launch; more fields are handled by looping launches
    constexpr int kThreads = 256;
    constexpr int kItems = 4;
    constexpr uint64_t kMagic = 0x6774623230306831ull; // "gtb200h1"
    constexpr int64_t kFlagBytes = 2 * 32 * 8;         // flags[parity][direction], uint64 epoch numbers
    constexpr int64_t kAlign = 256;

    struct region {
        int lo[3];
        int len[3];
        int64_t count; // 0: no such neighbour
    };

    struct blob {
        uint64_t magic;
        int64_t pid;
        int device;
        int pad;
        uint64_t arena; // raw device pointer (valid inside the exporting process)
        int64_t arena_bytes;
        int64_t recv_total;   // bytes of one parity of the receive area
        int64_t recv_off[27]; // byte offset of direction n inside one parity
        cudaIpcMemHandle_t ipc;
    };
    static_assert(sizeof(blob) <= GTB_HALO_BLOB_BYTES, "blob too large");

} // namespace

struct gtb_halo {
    gtb_halo_desc d[3];
    int nbr[27];
    int my_rank, max_fields, es, device;
    region send[27], recv[27];
    int64_t send_off[27], recv_off[27]; // byte offsets (max_fields sized slots)
    int64_t send_total, recv_total;
    char *send_arena; // local staging buffers
    char *arena;      // [flags][recv parity 0][recv parity 1], exported to the neighbours
    int64_t arena_bytes;
    char *peer_arena[27];
    int64_t peer_recv_off[27], peer_recv_total[27];
    void *opened[27];
    bool connected;
    uint64_t epoch; // starts at 1
    int *d_error;   // device flag set by a wait that timed out
    unsigned *d_counters; // per-direction block counters of the fused pack + signal launch
};

namespace {

    // common/halo_descriptor.hpp:90-164
    int lo_inside(const gtb_halo_desc &h, int e) { return e == 1 ? h.end - h.minus + 1 : h.begin; }
    int hi_inside(const gtb_halo_desc &h, int e) { return e == -1 ? h.begin + h.plus - 1 : h.end; }
    int lo_outside(const gtb_halo_desc &h, int e) { return e == 0 ? h.begin : (e == 1 ? h.end + 1 : h.begin - h.minus); }
    int hi_outside(const gtb_halo_desc &h, int e) { return e == 0 ? h.end : (e == 1 ? h.end + h.plus : h.begin - 1); }

    // Fused synchronisation of a transfer launch.  mode 1 (pack towards peers): the last block of every direction
    // raises the neighbour's flag once all blocks of that direction have made their stores visible system-wide.
    // mode 2 (unpack): every block first acquires its direction's flag.
    struct sync_args {
        uint64_t *flag[27];
        uint64_t epoch;
        unsigned *counters;      // [27], zero between launches
        unsigned blocks_per_dir; // gridDim.x * gridDim.z
        int *error;
        long long timeout_cycles;
        int mode; // 0 none, 1 signal after pack, 2 wait before unpack
    };

    struct xfer_args {
        region r[27];
        char *buf[27]; // message buffer per direction (local or NVLink-mapped)
        char *fields[kMaxFields];
        int64_t s1, s2; // element strides of storage dimensions 1 and 2
        int n_fields;
        sync_args sync;
    };

    __device__ __forceinline__ void wait_flag(const uint64_t *flag, uint64_t epoch, int *error, long long timeout, int n) {
        if (*reinterpret_cast<volatile int *>(error))
            return; // an earlier wait already timed out: do not stall every following exchange as well
        const long long t0 = clock64();
        for (;;) {
            uint64_t v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= epoch)
                break;
            if (clock64() - t0 > timeout) {
                atomicExch(error, 1 + n);
                break;
            }
            __nanosleep(100);
        }
    }

    // PACK: field -> buffer; else buffer -> field.
    template <class E, bool PACK>
    __global__ void __launch_bounds__(kThreads) xfer_kernel(const __grid_constant__ xfer_args a) {
        const int n = blockIdx.y, f = blockIdx.z;
        const region &r = a.r[n];
        if (r.count == 0)
            return;
        if (!PACK && a.sync.mode == 2 && a.sync.flag[n]) {
            if (threadIdx.x == 0)
                wait_flag(a.sync.flag[n], a.sync.epoch, a.sync.error, a.sync.timeout_cycles, n);
            __syncthreads();
        }
        const int64_t base = (int64_t)blockIdx.x * (kThreads * kItems) + threadIdx.x;
        if (base < r.count) {
            E *buf = reinterpret_cast<E *>(a.buf[n]) + (int64_t)f * r.count;
            E *fld = reinterpret_cast<E *>(a.fields[f]);
            const int l0 = r.len[0], l1 = r.len[1];
            int64_t idx[kItems];
            E v[kItems];
#pragma unroll
            for (int t = 0; t < kItems; ++t) {
                int64_t e = base + (int64_t)t * kThreads;
                if (e < r.count) {
                    int64_t q = e / l0;
                    int i0 = (int)(e - q * l0);
                    int64_t q2 = q / l1;
                    int i1 = (int)(q - q2 * l1);
                    idx[t] = (r.lo[0] + i0) + (r.lo[1] + i1) * a.s1 + (r.lo[2] + q2) * a.s2;
                    v[t] = PACK ? fld[idx[t]] : buf[e];
                }
            }
#pragma unroll
            for (int t = 0; t < kItems; ++t) {
                int64_t e = base + (int64_t)t * kThreads;
                if (e < r.count) {
                    if (PACK)
                        buf[e] = v[t];
                    else
                        fld[idx[t]] = v[t];
                }
            }
        }
        if (PACK && a.sync.mode == 1 && a.sync.flag[n]) {
            __threadfence_system(); // this thread's payload stores are visible to the peer before the block is counted
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned done = atomicAdd(&a.sync.counters[n], 1u) + 1u;
                if (done == a.sync.blocks_per_dir) {
                    a.sync.counters[n] = 0;
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.sync.flag[n]), "l"(a.sync.epoch)
                                 : "memory");
                }
            }
        }
    }

    struct push_args {
        const char *src[27];
        char *dst[27];
        int64_t bytes[27];
    };

    // send buffers -> neighbours' receive buffers (16-byte vectors; all slots are 256-byte aligned)
    __global__ void __launch_bounds__(kThreads) push_kernel(const __grid_constant__ push_args a) {
        const int n = blockIdx.y;
        const int64_t nvec = (a.bytes[n] + 15) / 16;
        const uint4 *s = reinterpret_cast<const uint4 *>(a.src[n]);
        uint4 *d = reinterpret_cast<uint4 *>(a.dst[n]);
        for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < nvec; e += (int64_t)gridDim.x * kThreads)
            d[e] = s[e];
    }

    struct signal_args {
        uint64_t *flag[27]; // neighbour's flag word for the message coming from this rank (nullptr: none)
        uint64_t epoch;
    };

    __global__ void signal_kernel(const __grid_constant__ signal_args a) {
        const int n = threadIdx.x;
        if (n < 27 && a.flag[n]) {
            __threadfence_system(); // order the payload stores of the previous kernels before the flag
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[n]), "l"(a.epoch) : "memory");
        }
    }

    struct wait_args {
        const uint64_t *flag[27]; // own flag word per direction (nullptr: nothing expected)
        uint64_t epoch;
        int *error;
        long long timeout_cycles;
    };

    __global__ void wait_kernel(const __grid_constant__ wait_args a) {
        const int n = threadIdx.x;
        if (n < 27 && a.flag[n])
            wait_flag(a.flag[n], a.epoch, a.error, a.timeout_cycles, n);
    }

    int n_of(int e0, int e1, int e2) { return (e0 + 1) + 3 * (e1 + 1) + 9 * (e2 + 1); }

    int64_t align_up(int64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

    constexpr long long kTimeoutCycles = 6000000000ll; // ~3 s at 2 GHz: a lost neighbour must not hang the GPU

    // sync_mode 0: plain copy; 1: raise the neighbours' flags when done (pack); 2: wait for the own flags first (unpack)
    template <bool PACK>
    int run_xfer(gtb_halo *h, void *const *fields, int n_fields, char *const bufs[27], cudaStream_t stream,
        int sync_mode = 0) {
        int64_t max_count = 0;
        xfer_args a;
        a.sync.mode = 0;
        a.sync.epoch = h->epoch;
        a.sync.counters = h->d_counters;
        a.sync.error = h->d_error;
        a.sync.timeout_cycles = kTimeoutCycles;
        for (int n = 0; n < 27; ++n)
            a.sync.flag[n] = nullptr;
        const region *regs = PACK ? h->send : h->recv;
        for (int n = 0; n < 27; ++n) {
            a.r[n] = regs[n];
            a.buf[n] = bufs[n];
            if (!bufs[n])
                a.r[n].count = 0;
            if (a.r[n].count > max_count)
                max_count = a.r[n].count;
        }
        if (max_count == 0)
            return GTB_OK;
        a.s1 = h->d[0].total;
        a.s2 = (int64_t)h->d[0].total * h->d[1].total;
        for (int f0 = 0; f0 < n_fields; f0 += kMaxFields) {
            int nf = n_fields - f0 < kMaxFields ? n_fields - f0 : kMaxFields;
            for (int f = 0; f < nf; ++f)
                a.fields[f] = static_cast<char *>(fields[f0 + f]);
            a.n_fields = nf;
            xfer_args b = a;
            for (int n = 0; n < 27; ++n)
                if (b.buf[n])
                    b.buf[n] += (int64_t)f0 * b.r[n].count * h->es;
            dim3 grid((unsigned)((max_count + kThreads * kItems - 1) / (kThreads * kItems)), 27, nf);
            const bool first = f0 == 0, last = f0 + nf >= n_fields;
            if ((sync_mode == 1 && last) || (sync_mode == 2 && first)) {
                b.sync.mode = sync_mode;
                b.sync.blocks_per_dir = grid.x * grid.z;
                for (int n = 0; n < 27; ++n) {
                    if (n == 13 || h->nbr[n] < 0 || b.r[n].count == 0)
                        continue;
                    if (sync_mode == 1) // the neighbour sees this rank in direction 26 - n
                        b.sync.flag[n] =
                            reinterpret_cast<uint64_t *>(h->peer_arena[n]) + (h->epoch & 1) * 32 + (26 - n);
                    else
                        b.sync.flag[n] = reinterpret_cast<uint64_t *>(h->arena) + (h->epoch & 1) * 32 + n;
                }
            }
            if (h->es == 8)
                xfer_kernel<uint64_t, PACK><<<grid, kThreads, 0, stream>>>(b);
            else
                xfer_kernel<uint32_t, PACK><<<grid, kThreads, 0, stream>>>(b);
            count_launch();
            int st = check_launch(PACK ? "halo pack" : "halo unpack");
            if (st)
                return st;
        }
        return GTB_OK;
    }

    int check_fields(const gtb_halo *h, void *const *fields, int n_fields, const char *who) {
        if (!h || !fields)
            return fail(GTB_ERR_ARG, "%s: null argument", who);
        if (n_fields < 0 || n_fields > h->max_fields)
            return fail(GTB_ERR_ARG, "%s: n_fields %d exceeds max_fields %d given to setup", who, n_fields,
                h->max_fields);
        for (int f = 0; f < n_fields; ++f)
            if (!fields[f])
                return fail(GTB_ERR_ARG, "%s: field %d is null", who, f);
        return GTB_OK;
    }

    char *recv_slot(const gtb_halo *h, int n, uint64_t epoch) {
        return h->arena + kFlagBytes + (int64_t)(epoch & 1) * h->recv_total + h->recv_off[n];
    }
    char *peer_slot(const gtb_halo *h, int n, uint64_t epoch) {
        return h->peer_arena[n] + kFlagBytes + (int64_t)(epoch & 1) * h->peer_recv_total[n] + h->peer_recv_off[n];
    }

    int signal(gtb_halo *h, cudaStream_t stream) {
        signal_args s;
        bool any = false;
        for (int n = 0; n < 27; ++n) {
            s.flag[n] = nullptr;
            if (n != 13 && h->nbr[n] >= 0 && h->peer_arena[n]) {
                // the neighbour sees this rank in direction 26 - n
                s.flag[n] = reinterpret_cast<uint64_t *>(h->peer_arena[n]) + (h->epoch & 1) * 32 + (26 - n);
                any = true;
            }
        }
        if (!any)
            return GTB_OK;
        s.epoch = h->epoch;
        signal_kernel<<<1, 32, 0, stream>>>(s);
        count_launch();
        return check_launch("halo signal");
    }

} // namespace

GTB_API int gtb_halo_create(const gtb_halo_desc desc[3], const int neighbour_rank[27], int my_rank, int max_fields,
    int elem_size, gtb_halo **out) {
    if (!desc || !neighbour_rank || !out)
        return fail(GTB_ERR_ARG, "gtb_halo_create: null argument");
    if (elem_size != 4 && elem_size != 8)
        return fail(GTB_ERR_ARG, "gtb_halo_create: elem_size %d not in {4,8}", elem_size);
    if (max_fields < 1)
        return fail(GTB_ERR_ARG, "gtb_halo_create: max_fields must be >= 1");
    for (int d = 0; d < 3; ++d) {
        const gtb_halo_desc &h = desc[d];
        if (h.minus < 0 || h.plus < 0 || h.begin < h.minus || h.end < h.begin || h.end + h.plus >= h.total)
            return fail(GTB_ERR_ARG,
                "gtb_halo_create: inconsistent halo descriptor %d (minus %d plus %d begin %d end %d total %d)", d,
                h.minus, h.plus, h.begin, h.end, h.total);
    }
    device_state *dv = dev();
    if (!dv)
        return GTB_ERR_CUDA;
    gtb_halo *h = new gtb_halo();
    memcpy(h->d, desc, sizeof(h->d));
    memcpy(h->nbr, neighbour_rank, sizeof(h->nbr));
    h->nbr[13] = -1;
    h->my_rank = my_rank;
    h->max_fields = max_fields;
    h->es = elem_size;
    h->device = dv->device;
    h->connected = false;
    h->epoch = 1;
    h->send_total = h->recv_total = 0;
    for (int e2 = -1; e2 <= 1; ++e2)
        for (int e1 = -1; e1 <= 1; ++e1)
            for (int e0 = -1; e0 <= 1; ++e0) {
                const int n = n_of(e0, e1, e2);
                const int e[3] = {e0, e1, e2};
                region &s = h->send[n], &r = h->recv[n];
                s.count = r.count = 1;
                for (int d = 0; d < 3; ++d) {
                    s.lo[d] = lo_inside(h->d[d], e[d]);
                    s.len[d] = hi_inside(h->d[d], e[d]) - s.lo[d] + 1;
                    r.lo[d] = lo_outside(h->d[d], e[d]);
                    r.len[d] = hi_outside(h->d[d], e[d]) - r.lo[d] + 1;
                    s.count *= s.len[d] > 0 ? s.len[d] : 0;
                    r.count *= r.len[d] > 0 ? r.len[d] : 0;
                }
                if (n == 13 || h->nbr[n] < 0)
                    s.count = r.count = 0;
                h->send_off[n] = h->send_total;
                h->recv_off[n] = h->recv_total;
                h->send_total += align_up(s.count * max_fields * elem_size);
                h->recv_total += align_up(r.count * max_fields * elem_size);
                h->peer_arena[n] = nullptr;
                h->opened[n] = nullptr;
                h->peer_recv_off[n] = h->peer_recv_total[n] = 0;
            }
    h->arena_bytes = kFlagBytes + 2 * h->recv_total + kAlign;
    h->send_arena = nullptr;
    h->arena = nullptr;
    h->d_error = nullptr;
    h->d_counters = nullptr;
    cudaError_t e = cudaMalloc(&h->send_arena, (size_t)(h->send_total + kAlign));
    if (e == cudaSuccess)
        e = cudaMalloc(&h->arena, (size_t)h->arena_bytes);
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_error, sizeof(int));
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_counters, 32 * sizeof(unsigned));
    if (e == cudaSuccess)
        e = cudaMemset(h->d_counters, 0, 32 * sizeof(unsigned));
    if (e == cudaSuccess)
        e = cudaMemset(h->arena, 0, (size_t)h->arena_bytes);
    if (e == cudaSuccess)
        e = cudaMemset(h->d_error, 0, sizeof(int));
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(h->send_arena), cudaFree(h->arena), cudaFree(h->d_error), cudaFree(h->d_counters);
        delete h;
        cuda_fail(e, "gtb_halo_create: buffer allocation");
        return GTB_ERR_ALLOC;
    }
    *out = h;
    return GTB_OK;
}

GTB_API int gtb_halo_destroy(gtb_halo *h) {
    if (!h)
        return GTB_OK;
    cudaDeviceSynchronize();
    for (int n = 0; n < 27; ++n)
        if (h->opened[n]) {
            bool shared = false; // one mapping may serve several directions
            for (int m = 0; m < n; ++m)
                shared = shared || h->opened[m] == h->opened[n];
            if (!shared)
                cudaIpcCloseMemHandle(h->opened[n]);
        }
    cudaFree(h->send_arena);
    cudaFree(h->arena);
    cudaFree(h->d_error);
    cudaFree(h->d_counters);
    delete h;
    return GTB_OK;
}

GTB_API int64_t gtb_halo_send_bytes(const gtb_halo *h, int n, int n_fields) {
    if (!h || n < 0 || n >= 27)
        return 0;
    return h->send[n].count * n_fields * h->es;
}

GTB_API int64_t gtb_halo_recv_bytes(const gtb_halo *h, int n, int n_fields) {
    if (!h || n < 0 || n >= 27)
        return 0;
    return h->recv[n].count * n_fields * h->es;
}

GTB_API void *gtb_halo_send_buffer(const gtb_halo *h, int n) {
    if (!h || n < 0 || n >= 27 || h->send[n].count == 0)
        return nullptr;
    return h->send_arena + h->send_off[n];
}

GTB_API void *gtb_halo_recv_buffer(const gtb_halo *h, int n) {
    if (!h || n < 0 || n >= 27 || h->recv[n].count == 0)
        return nullptr;
    return recv_slot(h, n, h->epoch);
}

GTB_API int gtb_halo_export(gtb_halo *h, void *out) {
    if (!h || !out)
        return fail(GTB_ERR_ARG, "gtb_halo_export: null argument");
    blob b;
    memset(&b, 0, sizeof(b));
    b.magic = kMagic;
    b.pid = (int64_t)getpid();
    b.device = h->device;
    b.arena = reinterpret_cast<uint64_t>(h->arena);
    b.arena_bytes = h->arena_bytes;
    b.recv_total = h->recv_total;
    memcpy(b.recv_off, h->recv_off, sizeof(b.recv_off));
    GTB_CUDA(cudaIpcGetMemHandle(&b.ipc, h->arena));
    memset(out, 0, GTB_HALO_BLOB_BYTES);
    memcpy(out, &b, sizeof(b));
    return GTB_OK;
}

GTB_API int gtb_halo_connect(gtb_halo *h, const void *const blobs[27]) {
    if (!h || !blobs)
        return fail(GTB_ERR_ARG, "gtb_halo_connect: null argument");
    GTB_CUDA(cudaSetDevice(h->device));
    for (int n = 0; n < 27; ++n) {
        if (n == 13 || h->nbr[n] < 0)
            continue;
        if (!blobs[n])
            return fail(GTB_ERR_ARG, "gtb_halo_connect: no blob for neighbour direction %d (rank %d)", n, h->nbr[n]);
        blob b;
        memcpy(&b, blobs[n], sizeof(b));
        if (b.magic != kMagic)
            return fail(GTB_ERR_ARG, "gtb_halo_connect: blob of direction %d is not a gtb_halo export", n);
        char *base = nullptr;
        if (b.pid == (int64_t)getpid()) {
            base = reinterpret_cast<char *>(b.arena);
            if (b.device != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled)
                    cudaGetLastError();
                else if (e != cudaSuccess)
                    return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            }
        } else {
            for (int m = 0; m < n && !base; ++m) // the same peer may be reached in several directions
                if (h->opened[m] && h->nbr[m] == h->nbr[n])
                    base = static_cast<char *>(h->opened[m]);
            if (!base) {
                void *p = nullptr;
                GTB_CUDA(cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
                base = static_cast<char *>(p);
            }
            h->opened[n] = base;
        }
        h->peer_arena[n] = base;
        h->peer_recv_total[n] = b.recv_total;
        h->peer_recv_off[n] = b.recv_off[26 - n]; // the neighbour receives this rank's message from direction -eta
    }
    h->connected = true;
    return GTB_OK;
}

GTB_API int gtb_halo_pack(gtb_halo *h, void *const *fields, int n_fields, void *stream) {
    int st = check_fields(h, fields, n_fields, "gtb_halo_pack");
    if (st)
        return st;
    char *bufs[27];
    for (int n = 0; n < 27; ++n)
        bufs[n] = h->send[n].count ? h->send_arena + h->send_off[n] : nullptr;
    return run_xfer<true>(h, fields, n_fields, bufs, as_stream(stream));
}

GTB_API int gtb_halo_pack_send(gtb_halo *h, void *const *fields, int n_fields, void *stream) {
    int st = check_fields(h, fields, n_fields, "gtb_halo_pack_send");
    if (st)
        return st;
    if (!h->connected)
        return fail(GTB_ERR_STATE, "gtb_halo_pack_send: gtb_halo_connect has not been called");
    char *bufs[27];
    for (int n = 0; n < 27; ++n)
        bufs[n] = h->send[n].count ? peer_slot(h, n, h->epoch) : nullptr;
    if (n_fields == 0)
        return signal(h, as_stream(stream));
    return run_xfer<true>(h, fields, n_fields, bufs, as_stream(stream), 1);
}

GTB_API int gtb_halo_send(gtb_halo *h, int n_fields, void *stream) {
    if (!h)
        return fail(GTB_ERR_ARG, "gtb_halo_send: null handle");
    if (!h->connected)
        return fail(GTB_ERR_STATE, "gtb_halo_send: gtb_halo_connect has not been called");
    if (n_fields < 0 || n_fields > h->max_fields)
        return fail(GTB_ERR_ARG, "gtb_halo_send: n_fields %d exceeds max_fields %d", n_fields, h->max_fields);
    push_args a;
    int64_t max_bytes = 0;
    for (int n = 0; n < 27; ++n) {
        a.src[n] = nullptr, a.dst[n] = nullptr, a.bytes[n] = 0;
        if (h->send[n].count) {
            a.src[n] = h->send_arena + h->send_off[n];
            a.dst[n] = peer_slot(h, n, h->epoch);
            a.bytes[n] = h->send[n].count * n_fields * h->es;
            if (a.bytes[n] > max_bytes)
                max_bytes = a.bytes[n];
        }
    }
    if (max_bytes > 0) {
        int64_t blocks = (max_bytes / 16 + kThreads * 8 - 1) / (kThreads * 8);
        if (blocks < 1)
            blocks = 1;
        if (blocks > 64)
            blocks = 64;
        push_kernel<<<dim3((unsigned)blocks, 27), kThreads, 0, as_stream(stream)>>>(a);
        count_launch();
        int st = check_launch("halo push");
        if (st)
            return st;
    }
    return signal(h, as_stream(stream));
}

GTB_API int gtb_halo_wait(gtb_halo *h, void *stream) {
    if (!h)
        return fail(GTB_ERR_ARG, "gtb_halo_wait: null handle");
    if (!h->connected)
        return fail(GTB_ERR_STATE, "gtb_halo_wait: gtb_halo_connect has not been called");
    wait_args w;
    bool any = false;
    for (int n = 0; n < 27; ++n) {
        w.flag[n] = nullptr;
        if (h->recv[n].count) {
            w.flag[n] = reinterpret_cast<const uint64_t *>(h->arena) + (h->epoch & 1) * 32 + n;
            any = true;
        }
    }
    if (!any)
        return GTB_OK;
    w.epoch = h->epoch;
    w.error = h->d_error;
    w.timeout_cycles = kTimeoutCycles;
    wait_kernel<<<1, 32, 0, as_stream(stream)>>>(w);
    count_launch();
    return check_launch("halo wait");
}

GTB_API int gtb_halo_unpack(gtb_halo *h, void *const *fields, int n_fields, void *stream) {
    int st = check_fields(h, fields, n_fields, "gtb_halo_unpack");
    if (st)
        return st;
    char *bufs[27];
    for (int n = 0; n < 27; ++n)
        bufs[n] = h->recv[n].count ? recv_slot(h, n, h->epoch) : nullptr;
    return run_xfer<false>(h, fields, n_fields, bufs, as_stream(stream));
}

GTB_API int gtb_halo_wait_unpack(gtb_halo *h, void *const *fields, int n_fields, void *stream) {
    int st = check_fields(h, fields, n_fields, "gtb_halo_wait_unpack");
    if (st)
        return st;
    if (!h->connected)
        return fail(GTB_ERR_STATE, "gtb_halo_wait_unpack: gtb_halo_connect has not been called");
    if (n_fields == 0)
        return gtb_halo_wait(h, stream);
    char *bufs[27];
    for (int n = 0; n < 27; ++n)
        bufs[n] = h->recv[n].count ? recv_slot(h, n, h->epoch) : nullptr;
    return run_xfer<false>(h, fields, n_fields, bufs, as_stream(stream), 2);
}

GTB_API int gtb_halo_exchange(gtb_halo *h, void *const *fields, int n_fields, void *stream) {
    int st = gtb_halo_pack_send(h, fields, n_fields, stream);
    if (st)
        return st;
    st = gtb_halo_wait_unpack(h, fields, n_fields, stream);
    if (st)
        return st;
    return gtb_halo_next_epoch(h);
}

GTB_API int gtb_halo_error(gtb_halo *h, int *code) {
    if (!h || !code)
        return fail(GTB_ERR_ARG, "gtb_halo_error: null argument");
    GTB_CUDA(cudaDeviceSynchronize());
    GTB_CUDA(cudaMemcpy(code, h->d_error, sizeof(int), cudaMemcpyDeviceToHost));
    return GTB_OK;
}

GTB_API int gtb_halo_next_epoch(gtb_halo *h) {
    if (!h)
        return fail(GTB_ERR_ARG, "gtb_halo_next_epoch: null handle");
    h->epoch += 1;
    return GTB_OK;
}
