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
#include "halo_device.cuh"

#include <unistd.h>

#include <map>
#include <mutex>
#include <vector>

using namespace gtb;
using namespace gtb::halo_dev;

namespace {

    constexpr int kBlocksPerSm = 5; // transfer kernels: register budget 64K / (5 * 256) = 51 per thread
    constexpr uint64_t kMagic = 0x6774623230306831ull; // "gtb200h1"
    constexpr int64_t kFlagBytes = 2 * 32 * 8;         // flags[parity][direction], uint64 epoch numbers
    constexpr int64_t kAlign = 256;

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
    region border[27];      // outside region of the directions that have no neighbour (domain border)
    const unsigned long long *gate_counter; // one-shot: the next wait_unpack launch waits for *gate_counter >= gate_value
    unsigned long long gate_value;
    unsigned long long *d_unpacked;         // epoch of the last completed unpack (device)
    int bc_kind;            // -1 none; GTB_BC_VALUE: unpack launches also fill the border regions with bc_bits
    uint64_t bc_bits;
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
    int *d_error;   // set by a wait that timed out: the device address of ...
    int *h_error;   // ... this word of mapped, pinned host memory (readable without synchronising)
    unsigned *d_counters; // block counter of the fused pack + signal launch
    unsigned long long *trace; // diagnosis (gtb_halo_set_trace): [epoch % 256][8] globaltimer stamps
};

namespace {

    // common/halo_descriptor.hpp:90-164
    int lo_inside(const gtb_halo_desc &h, int e) { return e == 1 ? h.end - h.minus + 1 : h.begin; }
    int hi_inside(const gtb_halo_desc &h, int e) { return e == -1 ? h.begin + h.plus - 1 : h.end; }
    int lo_outside(const gtb_halo_desc &h, int e) { return e == 0 ? h.begin : (e == 1 ? h.end + 1 : h.begin - h.minus); }
    int hi_outside(const gtb_halo_desc &h, int e) { return e == 0 ? h.end : (e == 1 ? h.end + h.plus : h.begin - 1); }

    struct xfer_args {
        seg_table t;
        char *fields[kMaxFields];
        int64_t s1, s2; // element strides of storage dimensions 1 and 2
        int n_fields;
        uint64_t fill_bits; // value written into segments without a buffer (boundary condition fused into the unpack)
        sync_args sync;
    };

    // The chunks of one segment table, walked with a grid stride.  PACK: field -> buffer; else buffer -> field.
    // WAIT: acquire a segment's flag before its first chunk.
    template <class E, bool PACK, bool WAIT>
    __device__ __forceinline__ void move_chunks(const seg_table &t, char *const *fields, int64_t s1, int64_t s2,
        const sync_args &sy, uint64_t fill_bits = 0) {
        const int total = t.chunk_start[t.n_seg];
        unsigned waited = 0, failed = 0; // segments whose flag this block has acquired / given up on (block-uniform)
        __shared__ int s_ok;
        if (WAIT && blockIdx.x == 0 && threadIdx.x == 0) // segments without elements: flags only (see fill_table)
            for (int f = 0; f < t.n_seg; ++f)
                if (t.flag[f] && t.chunk_start[f + 1] == t.chunk_start[f])
                    wait_flag(t.flag[f], sy.epoch, sy.error, sy.timeout_cycles, t.dir[f]);
        int s = 0;
        for (int ch = blockIdx.x; ch < total; ch += gridDim.x) {
            while (ch >= t.chunk_start[s + 1])
                ++s;
            if (WAIT && t.flag[s] && !((waited >> s) & 1u)) {
                if (threadIdx.x == 0) {
                    s_ok = wait_flag(t.flag[s], sy.epoch, sy.error, sy.timeout_cycles, t.dir[s]);
                    if (sy.trace)
                        atomicMax(sy.trace + 3, ptx::globaltimer());
                }
                __syncthreads();
                waited |= 1u << s;
                if (!s_ok)
                    failed |= 1u << s;
                __syncthreads();
            }
            if (WAIT && ((failed >> s) & 1u))
                continue; // the message never arrived: the halo keeps its old values, the error word is set
            const region &r = t.r[s];
            const int local = ch - t.chunk_start[s];
            const int f = local / t.chunks_per_field[s];
            const uint32_t first = (uint32_t)(local - f * t.chunks_per_field[s]) * (uint32_t)kChunk;
            const bool fill = !PACK && t.buf[s] == nullptr; // border segment: boundary value instead of a message
            E *buf = reinterpret_cast<E *>(t.buf[s]) + (int64_t)f * r.count;
            move_chunk<E, PACK>(r.lo[0], r.lo[1], r.lo[2], r.len[0], r.len[1], s1, s2, (uint32_t)r.count, first,
                reinterpret_cast<E *>(fields[f]), buf, fill, fill_bits);
        }
    }

    // After the pack: the last block of the launch raises the neighbours' flags once every block has made its
    // stores visible system-wide.
    __device__ __forceinline__ void signal_peers(const seg_table &t, const sync_args &sy, int *s_last) {
        // The block's payload stores are ordered before the barrier, the ONE system-scope fence of thread 0 after it
        // is cumulative over them (the pattern of a cooperative-groups grid sync).  Every thread fencing at system
        // scope -- 5 000 MEMBAR.SYS per pack -- slowed the stencil kernel that runs beside the exchange.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned done = atomicAdd(sy.counter, 1u) + 1u;
            *s_last = done == gridDim.x;
            if (*s_last)
                *sy.counter = 0;
        }
        __syncthreads();
        if (*s_last && threadIdx.x < t.n_seg && t.flag[threadIdx.x]) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(t.flag[threadIdx.x]), "l"(sy.epoch) : "memory");
        }
    }

    template <class E, bool PACK>
    __global__ void __launch_bounds__(kThreads, kBlocksPerSm) xfer_kernel(const __grid_constant__ xfer_args a) {
        __shared__ int s_last;
        if (a.sync.trace && threadIdx.x == 0 && blockIdx.x == 0)
            a.sync.trace[PACK ? 0 : 2] = ptx::globaltimer();
        if (!PACK && a.sync.gate) { // a stencil launch on another stream may still be reading the halos
            if (threadIdx.x == 0)
                ptx::gate_wait(a.sync.gate, a.sync.gate_value, a.sync.gate_timeouts);
            __syncthreads();
        }
        if (!PACK && a.sync.mode == 2)
            move_chunks<E, false, true>(a.t, a.fields, a.s1, a.s2, a.sync, a.fill_bits);
        else
            move_chunks<E, PACK, false>(a.t, a.fields, a.s1, a.s2, a.sync, a.fill_bits);
        if (PACK && a.sync.mode == 1)
            signal_peers(a.t, a.sync, &s_last);
        if (a.sync.trace && threadIdx.x == 0)
            atomicMax(a.sync.trace + (PACK ? 1 : 4), ptx::globaltimer());
        if (!PACK && a.sync.unpacked) { // publish "halos of epoch e are in place" for device-side gates
            __syncthreads();
            if (threadIdx.x == 0)
                __threadfence();
            if (threadIdx.x == 0 && atomicAdd(a.sync.counter2, 1u) + 1u == gridDim.x) {
                *a.sync.counter2 = 0;
                __threadfence();
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.sync.unpacked), "l"(a.sync.epoch) : "memory");
            }
        }
    }

    // The whole exchange in ONE launch: pack into the neighbours' receive buffers, raise their flags, then acquire
    // the own flags and unpack.  The grid is small (it fits beside a persistent stencil kernel), so every block is
    // resident when it starts to wait; a block that waits only depends on the PEERS' pack phases, never on a block of
    // its own launch that has not started (those only delay the peers' flags until the scheduler places them).
    template <class E>
    __global__ void __launch_bounds__(kThreads, kBlocksPerSm) exchange_kernel(const __grid_constant__ exchange_args a) {
        __shared__ int s_last;
        move_chunks<E, true, false>(a.snd, a.fields, a.s1, a.s2, a.sync);
        signal_peers(a.snd, a.sync, &s_last);
        move_chunks<E, false, true>(a.rcv, a.fields, a.s1, a.s2, a.sync, a.fill_bits);
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
        unsigned long long *trace; // diagnosis: [5] = flags raised
    };

    __global__ void signal_kernel(const __grid_constant__ signal_args a) {
        const int n = threadIdx.x;
        if (a.trace && n == 0)
            a.trace[5] = ptx::globaltimer();
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

    // Clock cycles a device-side wait for a neighbour's message may take (option halo.timeout_ms, default 60 s; 0: wait
    // for ever like MPI_Wait).  Long, because a neighbour that is merely late -- load imbalance, I/O, a first-call
    // JIT, a debugger -- is not an error; finite by default, because a GPU that spins for ever cannot be recovered.
    long long timeout_cycles() {
        const long long ms = opts().halo_timeout_ms;
        return ms <= 0 ? 0 : ms * 2000000ll; // ~2 GHz
    }

    // Segment table of the active directions.  sync_mode 1: flags to raise at the neighbours; 2: own flags to wait for.
    void fill_table(seg_table &t, const gtb_halo *h, bool pack, char *const bufs[27], int nf, int64_t field_offset,
        int sync_mode, int chunk = kChunk) {
        const region *regs = pack ? h->send : h->recv;
        t.n_seg = 0;
        t.chunk_start[0] = 0;
        for (int n = 0; n < 27; ++n) {
            if (n == 13)
                continue;
            // A direction that carries payload only the OTHER way (asymmetric halos: minus = 1, plus = 0) still
            // exchanges flags both ways -- a segment without elements: seeing a neighbour's epoch e + 1 is what proves
            // that it has unpacked epoch e and its double-buffered slot may be overwritten (MPI needs no such thing).
            const bool flag_only = sync_mode != 0 && h->nbr[n] >= 0 && regs[n].count == 0 &&
                                   (pack ? h->recv[n].count : h->send[n].count) != 0;
            if (!flag_only && (!bufs[n] || regs[n].count == 0))
                continue;
            const int sg = t.n_seg++;
            t.dir[sg] = n;
            t.r[sg] = regs[n];
            t.buf[sg] = flag_only ? nullptr : bufs[n] + field_offset * regs[n].count * h->es;
            t.flag[sg] = nullptr;
            if (h->nbr[n] >= 0) {
                if (sync_mode == 1) // the neighbour sees this rank in direction 26 - n
                    t.flag[sg] = reinterpret_cast<uint64_t *>(h->peer_arena[n]) + (h->epoch & 1) * 32 + (26 - n);
                else if (sync_mode == 2)
                    t.flag[sg] = reinterpret_cast<uint64_t *>(h->arena) + (h->epoch & 1) * 32 + n;
            }
            t.chunks_per_field[sg] = (int)((regs[n].count + chunk - 1) / chunk);
            t.chunk_start[sg + 1] = t.chunk_start[sg] + t.chunks_per_field[sg] * nf;
        }
        if (!pack && h->bc_kind == GTB_BC_VALUE) // distributed_boundaries.hpp: value condition where there is no neighbour
            for (int n = 0; n < 27 && t.n_seg < kMaxSeg; ++n) {
                if (h->border[n].count == 0)
                    continue;
                const int sg = t.n_seg++;
                t.dir[sg] = n;
                t.r[sg] = h->border[n];
                t.buf[sg] = nullptr;
                t.flag[sg] = nullptr;
                t.chunks_per_field[sg] = (int)((h->border[n].count + chunk - 1) / chunk);
                t.chunk_start[sg + 1] = t.chunk_start[sg] + t.chunks_per_field[sg] * nf;
            }
    }

    void fill_sync(sync_args &sy, const gtb_halo *h, int mode) {
        sy.mode = mode;
        sy.epoch = h->epoch;
        sy.counter = h->d_counters;
        sy.error = h->d_error;
        sy.timeout_cycles = timeout_cycles();
        sy.gate = nullptr;
        sy.gate_value = 0;
        sy.gate_timeouts = nullptr;
        sy.unpacked = nullptr;
        sy.counter2 = h->d_counters + 1;
        sy.trace = h->trace ? h->trace + (h->epoch % 256) * 8 : nullptr;
    }

    // Blocks of a transfer launch.  Beside a persistent stencil kernel (reserve_sms > 0) the grid is what the reserved
    // SMs hold at once (kBlocksPerSm each): every block is resident from the start and the launch does not have to
    // wait for stencil CTAs to retire (the device timeline of round 2, profiles/r02_exchange_timeline.txt, showed the
    // tail blocks of a one-block-per-SM grid doing exactly that).  Standing alone: one block per SM.  Option
    // halo.max_blocks overrides both.
    int xfer_grid(int chunks) {
        device_state *dv = dev();
        int cap = dv ? dv->sm_count : 128;
        if (opts().reserve_sms > 0)
            cap = opts().reserve_sms * kBlocksPerSm;
        if (opts().halo_max_blocks > 0)
            cap = opts().halo_max_blocks;
        return chunks < 1 ? 1 : (chunks > cap ? cap : chunks);
    }

    // Rewrites a transfer in units of 16-byte vectors when every row piece of every segment is a whole number of aligned
    // vectors (hori_diff's halo of 2 doubles: one vector per row of an I face, 128 per row of a J face): half (fp64) or a
    // quarter (fp32) of the elements, chunks and instructions.  The byte order of a message does not change, so a
    // vectorised pack and a scalar unpack of the same message agree.
    bool vectorize(xfer_args &b, int es) {
        const int V = 16 / es;
        if (b.s1 % V || b.s2 % V)
            return false;
        for (int f = 0; f < b.n_fields; ++f)
            if (reinterpret_cast<uintptr_t>(b.fields[f]) % 16)
                return false;
        for (int sg = 0; sg < b.t.n_seg; ++sg) {
            const region &r = b.t.r[sg];
            if (r.lo[0] % V || r.len[0] % V || reinterpret_cast<uintptr_t>(b.t.buf[sg]) % 16)
                return false;
        }
        b.s1 /= V;
        b.s2 /= V;
        for (int sg = 0; sg < b.t.n_seg; ++sg) {
            region &r = b.t.r[sg];
            r.lo[0] /= V;
            r.len[0] /= V;
            r.count /= V;
            b.t.chunks_per_field[sg] = (int)((r.count + kChunk - 1) / kChunk);
            b.t.chunk_start[sg + 1] = b.t.chunk_start[sg] + b.t.chunks_per_field[sg] * b.n_fields;
        }
        return true;
    }

    // sync_mode 0: plain copy; 1: raise the neighbours' flags when done (pack); 2: wait for the own flags first (unpack)
    template <bool PACK>
    int run_xfer(gtb_halo *h, void *const *fields, int n_fields, char *const bufs[27], cudaStream_t stream,
        int sync_mode = 0) {
        for (int f0 = 0; f0 < n_fields; f0 += kMaxFields) {
            const int nf = n_fields - f0 < kMaxFields ? n_fields - f0 : kMaxFields;
            const bool first = f0 == 0, last = f0 + nf >= n_fields;
            const int mode = (sync_mode == 1 && last) || (sync_mode == 2 && first) ? sync_mode : 0;
            xfer_args b;
            fill_table(b.t, h, PACK, bufs, nf, f0, mode);
            if (b.t.n_seg == 0)
                return GTB_OK;
            b.fill_bits = h->bc_bits;
            fill_sync(b.sync, h, mode);
            if (!PACK && sync_mode == 2) {
                if (first && h->gate_counter) {
                    b.sync.gate = h->gate_counter;
                    b.sync.gate_value = h->gate_value;
                    b.sync.gate_timeouts = gate_timeout_counter();
                    h->gate_counter = nullptr; // one-shot
                }
                if (last)
                    b.sync.unpacked = h->d_unpacked;
            }
            b.s1 = h->d[0].total;
            b.s2 = (int64_t)h->d[0].total * h->d[1].total;
            for (int f = 0; f < nf; ++f)
                b.fields[f] = static_cast<char *>(fields[f0 + f]);
            b.n_fields = nf;
            const bool vec = opts().halo_vec && vectorize(b, h->es);
            const int grid = xfer_grid(b.t.chunk_start[b.t.n_seg]);
            if (vec)
                xfer_kernel<uint4, PACK><<<grid, kThreads, 0, stream>>>(b);
            else if (h->es == 8)
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
            if (n != 13 && h->nbr[n] >= 0 && h->peer_arena[n] && (h->send[n].count || h->recv[n].count)) {
                // the neighbour sees this rank in direction 26 - n
                s.flag[n] = reinterpret_cast<uint64_t *>(h->peer_arena[n]) + (h->epoch & 1) * 32 + (26 - n);
                any = true;
            }
        }
        if (!any)
            return GTB_OK;
        s.epoch = h->epoch;
        s.trace = h->trace ? h->trace + (h->epoch % 256) * 8 : nullptr;
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
    h->bc_kind = -1;
    h->gate_counter = nullptr;
    h->gate_value = 0;
    h->d_unpacked = nullptr;
    h->trace = nullptr;
    h->bc_bits = 0;
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
                h->border[n] = r;
                if (n == 13 || h->nbr[n] >= 0)
                    h->border[n].count = 0;
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
    h->h_error = nullptr;
    h->d_counters = nullptr;
    cudaError_t e = cudaMalloc(&h->send_arena, (size_t)(h->send_total + kAlign));
    if (e == cudaSuccess)
        e = cudaMalloc(&h->arena, (size_t)h->arena_bytes);
    if (e == cudaSuccess)
        e = cudaHostAlloc(&h->h_error, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *h->h_error = 0;
        e = cudaHostGetDevicePointer(&h->d_error, h->h_error, 0);
    }
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_counters, 32 * sizeof(unsigned));
    if (e == cudaSuccess)
        e = cudaMemset(h->d_counters, 0, 32 * sizeof(unsigned));
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_unpacked, sizeof(unsigned long long));
    if (e == cudaSuccess)
        e = cudaMemset(h->d_unpacked, 0, sizeof(unsigned long long));
    if (e == cudaSuccess)
        e = cudaMemset(h->arena, 0, (size_t)h->arena_bytes);
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(h->send_arena), cudaFree(h->arena), cudaFreeHost(h->h_error), cudaFree(h->d_counters), cudaFree(h->d_unpacked);
        delete h;
        cuda_fail(e, "gtb_halo_create: buffer allocation");
        return GTB_ERR_ALLOC;
    }
    *out = h;
    return GTB_OK;
}

namespace {
    void free_generic_tables(const gtb_halo *h);
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
    free_generic_tables(h);
    cudaFree(h->send_arena);
    cudaFree(h->arena);
    cudaFreeHost(h->h_error);
    cudaFree(h->d_counters);
    cudaFree(h->d_unpacked);
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
    if (opts().halo_dma) {
        // The NVLink leg on the COPY ENGINES: one asynchronous peer copy per neighbour from the packed send buffer into
        // the neighbour's receive buffer.  An SM sustains only ~10 GB/s of peer stores (profiles/r02_exchange_timeline.txt:
        // the 2 MB of an 8-neighbour exchange took 33 us on the four SMs a persistent stencil can spare), the copy
        // engines move them at link speed without touching an SM.  The flags follow in stream order.
        for (int n = 0; n < 27; ++n)
            if (h->send[n].count)
                GTB_CUDA(cudaMemcpyAsync(peer_slot(h, n, h->epoch), h->send_arena + h->send_off[n],
                    (size_t)(h->send[n].count * n_fields * h->es), cudaMemcpyDefault, as_stream(stream)));
        return signal(h, as_stream(stream));
    }
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
        if (n != 13 && h->nbr[n] >= 0 && (h->recv[n].count || h->send[n].count)) { // flags travel both ways
            w.flag[n] = reinterpret_cast<const uint64_t *>(h->arena) + (h->epoch & 1) * 32 + n;
            any = true;
        }
    }
    if (!any)
        return GTB_OK;
    w.epoch = h->epoch;
    w.error = h->d_error;
    w.timeout_cycles = timeout_cycles();
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
    int st = check_fields(h, fields, n_fields, "gtb_halo_exchange");
    if (st)
        return st;
    if (!h->connected)
        return fail(GTB_ERR_STATE, "gtb_halo_exchange: gtb_halo_connect has not been called");
    // (the one-launch kernel takes no device-side gate and does not publish the unpacked epoch)
    if (n_fields >= 1 && n_fields <= kMaxFields && opts().halo_fused && !h->gate_counter) { // one launch: pack, signal, wait, unpack
        char *sbufs[27], *rbufs[27];
        for (int n = 0; n < 27; ++n) {
            sbufs[n] = h->send[n].count ? peer_slot(h, n, h->epoch) : nullptr;
            rbufs[n] = h->recv[n].count ? recv_slot(h, n, h->epoch) : nullptr;
        }
        exchange_args a;
        fill_table(a.snd, h, true, sbufs, n_fields, 0, 1);
        fill_table(a.rcv, h, false, rbufs, n_fields, 0, 2);
        if (a.snd.n_seg == 0 && a.rcv.n_seg == 0)
            return gtb_halo_next_epoch(h);
        fill_sync(a.sync, h, 1);
        a.fill_bits = h->bc_bits;
        a.s1 = h->d[0].total;
        a.s2 = (int64_t)h->d[0].total * h->d[1].total;
        for (int f = 0; f < n_fields; ++f)
            a.fields[f] = static_cast<char *>(fields[f]);
        a.n_fields = n_fields;
        const int cs = a.snd.chunk_start[a.snd.n_seg], cr = a.rcv.chunk_start[a.rcv.n_seg];
        int grid = xfer_grid(cs > cr ? cs : cr);
        if (grid > 64)
            grid = 64; // every block must be resident while it waits for the peers
        if (h->es == 8)
            exchange_kernel<uint64_t><<<grid, kThreads, 0, as_stream(stream)>>>(a);
        else
            exchange_kernel<uint32_t><<<grid, kThreads, 0, as_stream(stream)>>>(a);
        count_launch();
        st = check_launch("halo exchange");
        if (st)
            return st;
        return gtb_halo_next_epoch(h);
    }
    if (opts().halo_dma) { // pack locally, copy engines over NVLink, flags (option halo.dma, see gtb_halo_send)
        st = gtb_halo_pack(h, fields, n_fields, stream);
        if (!st)
            st = gtb_halo_send(h, n_fields, stream);
    } else {
        st = gtb_halo_pack_send(h, fields, n_fields, stream);
    }
    if (st)
        return st;
    st = gtb_halo_wait_unpack(h, fields, n_fields, stream);
    if (st)
        return st;
    return gtb_halo_next_epoch(h);
}

// ------------------------------------------------------------------------------------------- generic exchange
// halo_exchange_generic (gcl/halo_exchange.hpp:335-513): every field has its own halo descriptors.  The reference
// concatenates the fields into one message per neighbour (descriptor_generic_manual.hpp:370-796) with per-field
// kernels; here the (direction, field) pieces of ALL fields form one flat segment list in device memory that one
// launch walks -- one launch packs and pushes everything, one launch waits and unpacks everything.
namespace {
    struct gseg {
        int lo[3], len[3];
        int dir;          // direction number 0..26
        int chunk_start;  // first chunk of this segment in the flat chunk list
        int64_t count;    // words
        int64_t s1, s2;   // word strides of storage dimensions 1 and 2 of this field
        char *fld;        // storage element (0,0,0) of the field
        char *buf;        // where this piece lives in the message of its direction
    };

    struct gx_args {
        const gseg *segs;
        int n_seg, n_chunks;
        uint64_t *flag[27]; // pack: flags to raise at the neighbours; unpack: own flags to wait for
        sync_args sync;
    };

    template <class E, bool PACK>
    __global__ void __launch_bounds__(kThreads, kBlocksPerSm) generic_xfer_kernel(const __grid_constant__ gx_args a) {
        __shared__ int s_last;
        if (a.sync.trace && threadIdx.x == 0 && blockIdx.x == 0)
            a.sync.trace[PACK ? 0 : 2] = ptx::globaltimer();
        unsigned waited = 0, failed = 0; // directions whose flag this block has acquired / given up on
        if (!PACK && blockIdx.x == 0 && threadIdx.x == 0) { // directions without a segment of their own: flags only
            unsigned with_seg = 0;
            for (int f = 0; f < a.n_seg; ++f)
                with_seg |= 1u << a.segs[f].dir;
            for (int n = 0; n < 27; ++n)
                if (a.flag[n] && !((with_seg >> n) & 1u))
                    wait_flag(a.flag[n], a.sync.epoch, a.sync.error, a.sync.timeout_cycles, n);
        }
        int s = 0;
        for (int ch = blockIdx.x; ch < a.n_chunks; ch += gridDim.x) {
            while (s + 1 < a.n_seg && ch >= a.segs[s + 1].chunk_start)
                ++s;
            const gseg g = a.segs[s];
            if (!PACK && a.flag[g.dir] && !((waited >> g.dir) & 1u)) {
                if (threadIdx.x == 0)
                    s_last = wait_flag(a.flag[g.dir], a.sync.epoch, a.sync.error, a.sync.timeout_cycles, g.dir);
                __syncthreads();
                waited |= 1u << g.dir;
                if (!s_last)
                    failed |= 1u << g.dir;
                __syncthreads();
            }
            if (!PACK && ((failed >> g.dir) & 1u))
                continue;
            move_chunk<E, PACK>(g.lo[0], g.lo[1], g.lo[2], g.len[0], g.len[1], g.s1, g.s2, (uint32_t)g.count,
                (uint32_t)(ch - g.chunk_start) * (uint32_t)kChunk, reinterpret_cast<E *>(g.fld),
                reinterpret_cast<E *>(g.buf), false, 0);
        }
        if (PACK) { // the last block raises the neighbours' flags once every block's stores are visible system-wide
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence_system(); // cumulative over the block's stores (see signal_peers)
                const unsigned done = atomicAdd(a.sync.counter, 1u) + 1u;
                s_last = done == gridDim.x;
                if (s_last)
                    *a.sync.counter = 0;
            }
            __syncthreads();
            if (s_last && threadIdx.x < 27 && a.flag[threadIdx.x]) {
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[threadIdx.x]), "l"(a.sync.epoch) : "memory");
            }
        }
        if (a.sync.trace && threadIdx.x == 0)
            atomicMax(a.sync.trace + (PACK ? 1 : 4), ptx::globaltimer());
    }

    // Device copies of segment lists, keyed by the bytes that determine them: a time loop that exchanges the same
    // fields every step builds and uploads its four tables (pack / unpack x epoch parity) once.
    struct gtable {
        std::vector<char> key;
        gseg *dev;
        int n_seg, n_chunks;
    };
    std::mutex g_gtable_mutex;
    std::map<const gtb_halo *, std::vector<gtable>> g_gtables;

    template <bool PACK>
    int run_generic(gtb_halo *h, const gtb_halo_field *fields, int n_fields, cudaStream_t stream, const char *who) {
        if (!h || (!fields && n_fields))
            return fail(GTB_ERR_ARG, "%s: null argument", who);
        if (!h->connected)
            return fail(GTB_ERR_STATE, "%s: gtb_halo_connect has not been called", who);
        if (n_fields < 0)
            return fail(GTB_ERR_ARG, "%s: negative field count", who);
        const int parity = (int)(h->epoch & 1);
        std::vector<char> key(sizeof(int) * 2 + sizeof(gtb_halo_field) * (size_t)n_fields);
        const int tag[2] = {PACK ? 1 : 0, parity};
        memcpy(key.data(), tag, sizeof(tag));
        if (n_fields)
            memcpy(key.data() + sizeof(tag), fields, sizeof(gtb_halo_field) * (size_t)n_fields);
        gtable *tab = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_gtable_mutex);
            std::vector<gtable> &tabs = g_gtables[h];
            for (gtable &t : tabs)
                if (t.key == key)
                    tab = &t;
            if (!tab) {
                std::vector<gseg> segs;
                int chunks = 0;
                for (int n = 0; n < 27; ++n) {
                    if (n == 13 || h->nbr[n] < 0)
                        continue;
                    const int e[3] = {n % 3 - 1, (n / 3) % 3 - 1, n / 9 - 1};
                    char *base = PACK ? (h->send[n].count ? peer_slot(h, n, h->epoch) : nullptr)
                                      : (h->recv[n].count ? recv_slot(h, n, h->epoch) : nullptr);
                    const int64_t capacity = (PACK ? h->send[n].count : h->recv[n].count) * h->max_fields * h->es;
                    int64_t offset = 0;
                    for (int f = 0; f < n_fields; ++f) {
                        const gtb_halo_desc *d = fields[f].desc;
                        gseg g;
                        g.count = 1;
                        for (int x = 0; x < 3; ++x) {
                            if (d[x].minus < 0 || d[x].plus < 0 || d[x].begin < d[x].minus || d[x].end < d[x].begin ||
                                d[x].end + d[x].plus >= d[x].total)
                                return fail(GTB_ERR_ARG, "%s: inconsistent halo descriptor of field %d, dimension %d", who, f, x);
                            g.lo[x] = PACK ? lo_inside(d[x], e[x]) : lo_outside(d[x], e[x]);
                            g.len[x] = (PACK ? hi_inside(d[x], e[x]) : hi_outside(d[x], e[x])) - g.lo[x] + 1;
                            g.count *= g.len[x] > 0 ? g.len[x] : 0;
                        }
                        if (g.count == 0)
                            continue;
                        if (!fields[f].ptr)
                            return fail(GTB_ERR_ARG, "%s: field %d is null", who, f);
                        if (!base || offset + g.count * h->es > capacity)
                            return fail(GTB_ERR_ARG,
                                "%s: the message for direction %d does not fit the buffers sized by the halo example at "
                                "setup (%lld bytes)", who, n, (long long)capacity);
                        g.dir = n;
                        g.chunk_start = chunks;
                        g.s1 = d[0].total;
                        g.s2 = (int64_t)d[0].total * d[1].total;
                        g.fld = static_cast<char *>(fields[f].ptr);
                        g.buf = base + offset;
                        offset += g.count * h->es;
                        chunks += (int)((g.count + kChunk - 1) / kChunk);
                        segs.push_back(g);
                    }
                }
                gtable t;
                t.key = key;
                t.dev = nullptr;
                t.n_seg = (int)segs.size();
                t.n_chunks = chunks;
                if (!segs.empty()) {
                    GTB_CUDA(cudaMalloc(&t.dev, segs.size() * sizeof(gseg)));
                    GTB_CUDA(cudaMemcpy(t.dev, segs.data(), segs.size() * sizeof(gseg), cudaMemcpyHostToDevice));
                }
                if (tabs.size() >= 64) { // bounded: a program that keeps changing its field lists recycles the oldest
                    cudaStreamSynchronize(stream);
                    cudaFree(tabs.front().dev);
                    tabs.erase(tabs.begin());
                }
                tabs.push_back(std::move(t));
                tab = &tabs.back();
            }
        }
        gx_args a;
        a.segs = tab->dev;
        a.n_seg = tab->n_seg;
        a.n_chunks = tab->n_chunks;
        bool any_flag = false;
        for (int n = 0; n < 27; ++n) {
            a.flag[n] = nullptr;
            if (n == 13 || h->nbr[n] < 0)
                continue;
            const bool active = h->send[n].count || h->recv[n].count; // flags travel both ways (see fill_table)
            if (PACK && active && h->peer_arena[n])
                a.flag[n] = reinterpret_cast<uint64_t *>(h->peer_arena[n]) + (h->epoch & 1) * 32 + (26 - n);
            else if (!PACK && active)
                a.flag[n] = reinterpret_cast<uint64_t *>(h->arena) + (h->epoch & 1) * 32 + n;
            any_flag = any_flag || a.flag[n];
        }
        if (!any_flag && a.n_chunks == 0)
            return GTB_OK;
        fill_sync(a.sync, h, PACK ? 1 : 2);
        const int grid = xfer_grid(a.n_chunks);
        if (h->es == 8)
            generic_xfer_kernel<uint64_t, PACK><<<grid, kThreads, 0, stream>>>(a);
        else
            generic_xfer_kernel<uint32_t, PACK><<<grid, kThreads, 0, stream>>>(a);
        count_launch();
        return check_launch(who);
    }
} // namespace

namespace {
    void free_generic_tables(const gtb_halo *h) {
        std::lock_guard<std::mutex> lock(g_gtable_mutex);
        auto it = g_gtables.find(h);
        if (it == g_gtables.end())
            return;
        for (gtable &t : it->second)
            cudaFree(t.dev);
        g_gtables.erase(it);
    }
} // namespace

GTB_API int gtb_halo_generic_pack_send(gtb_halo *h, const gtb_halo_field *fields, int n_fields, void *stream) {
    return run_generic<true>(h, fields, n_fields, as_stream(stream), "gtb_halo_generic_pack_send");
}

GTB_API int gtb_halo_generic_wait_unpack(gtb_halo *h, const gtb_halo_field *fields, int n_fields, void *stream) {
    return run_generic<false>(h, fields, n_fields, as_stream(stream), "gtb_halo_generic_wait_unpack");
}

GTB_API int gtb_halo_error(gtb_halo *h, int *code) {
    if (!h || !code)
        return fail(GTB_ERR_ARG, "gtb_halo_error: null argument");
    GTB_CUDA(cudaDeviceSynchronize());
    *code = *reinterpret_cast<volatile int *>(h->h_error);
    return GTB_OK;
}

GTB_API int gtb_halo_poll_error(gtb_halo *h, int *code) {
    if (!h || !code)
        return fail(GTB_ERR_ARG, "gtb_halo_poll_error: null argument");
    *code = *reinterpret_cast<volatile int *>(h->h_error);
    return GTB_OK;
}

GTB_API int gtb_halo_set_trace(gtb_halo *h, void *device_u64_256x8) {
    if (!h)
        return fail(GTB_ERR_ARG, "gtb_halo_set_trace: null handle");
    h->trace = static_cast<unsigned long long *>(device_u64_256x8);
    return GTB_OK;
}

namespace {
    __global__ void stamp_kernel(unsigned long long *dst) { *dst = ptx::globaltimer(); }
} // namespace

GTB_API int gtb_stamp(void *device_u64, void *stream) {
    if (!device_u64)
        return fail(GTB_ERR_ARG, "gtb_stamp: null pointer");
    stamp_kernel<<<1, 1, 0, as_stream(stream)>>>(static_cast<unsigned long long *>(device_u64));
    return check_launch("stamp");
}

GTB_API int gtb_halo_next_epoch(gtb_halo *h) {
    if (!h)
        return fail(GTB_ERR_ARG, "gtb_halo_next_epoch: null handle");
    h->epoch += 1;
    return GTB_OK;
}

// ------------------------------------------------------------------------------------------- boundary conditions
// boundaries/boundary.hpp:57-72 with the predefined conditions of zero.hpp / value.hpp / copy.hpp.  The reference's GPU
// path (apply_gpu.hpp:236-313) launches one kernel per call with a 3-d thread block per direction group; here the
// outside regions of all selected directions and all fields are one flat chunk list walked by a small grid.
namespace {
    struct bc_args {
        seg_table t;
        char *fields[kMaxFields];
        const char *src; // copy_boundary: the last field
        int64_t s1, s2;
        int n_fields;
        uint64_t fill_bits;
    };

    template <class E, bool COPY>
    __global__ void __launch_bounds__(kThreads) bc_kernel(const __grid_constant__ bc_args a) {
        const seg_table &t = a.t;
        const int total = t.chunk_start[t.n_seg];
        int s = 0;
        for (int ch = blockIdx.x; ch < total; ch += gridDim.x) {
            while (ch >= t.chunk_start[s + 1])
                ++s;
            const region &r = t.r[s];
            const int local = ch - t.chunk_start[s];
            const int f = local / t.chunks_per_field[s];
            const int64_t base = (int64_t)(local - f * t.chunks_per_field[s]) * kChunk + threadIdx.x;
            E *fld = reinterpret_cast<E *>(a.fields[f]);
            const E *src = reinterpret_cast<const E *>(a.src);
            const int l0 = r.len[0], l1 = r.len[1];
#pragma unroll
            for (int it = 0; it < kItems; ++it) {
                int64_t e = base + (int64_t)it * kThreads;
                if (e < r.count) {
                    int64_t q = e / l0;
                    int i0 = (int)(e - q * l0);
                    int64_t q2 = q / l1;
                    int i1 = (int)(q - q2 * l1);
                    const int64_t idx = (r.lo[0] + i0) + (r.lo[1] + i1) * a.s1 + (r.lo[2] + q2) * a.s2;
                    fld[idx] = COPY ? src[idx] : (E)a.fill_bits;
                }
            }
        }
    }

    uint64_t value_bits(double value, int es) {
        uint64_t bits = 0;
        if (es == 8)
            memcpy(&bits, &value, 8);
        else {
            float f = (float)value;
            uint32_t b32;
            memcpy(&b32, &f, 4);
            bits = (uint64_t)b32 | ((uint64_t)b32 << 32); // the element pattern repeated to 64 bits (16-byte vector fills)
        }
        return bits;
    }
} // namespace

GTB_API int gtb_boundary_apply(const gtb_halo_desc desc[3], const int direction_mask[27], int kind, double value,
    void *const *fields, int n_fields, int elem_size, void *stream) {
    if (!desc || !fields)
        return fail(GTB_ERR_ARG, "gtb_boundary_apply: null argument");
    if (elem_size != 4 && elem_size != 8)
        return fail(GTB_ERR_ARG, "gtb_boundary_apply: elem_size %d not in {4,8}", elem_size);
    if (kind != GTB_BC_VALUE && kind != GTB_BC_COPY)
        return fail(GTB_ERR_ARG, "gtb_boundary_apply: unknown condition %d", kind);
    const int n_dst = kind == GTB_BC_COPY ? n_fields - 1 : n_fields;
    if (n_dst < 1 || n_dst > kMaxFields)
        return fail(GTB_ERR_ARG, "gtb_boundary_apply: %d destination fields (1..%d; copy_boundary needs a source as the last field)",
            n_dst, kMaxFields);
    for (int f = 0; f < n_fields; ++f)
        if (!fields[f])
            return fail(GTB_ERR_ARG, "gtb_boundary_apply: field %d is null", f);
    for (int d = 0; d < 3; ++d) {
        const gtb_halo_desc &h = desc[d];
        if (h.minus < 0 || h.plus < 0 || h.begin < h.minus || h.end < h.begin || h.end + h.plus >= h.total)
            return fail(GTB_ERR_ARG, "gtb_boundary_apply: inconsistent halo descriptor %d", d);
    }
    if (!dev())
        return GTB_ERR_CUDA;
    bc_args a;
    a.t.n_seg = 0;
    a.t.chunk_start[0] = 0;
    for (int e2 = -1; e2 <= 1; ++e2)
        for (int e1 = -1; e1 <= 1; ++e1)
            for (int e0 = -1; e0 <= 1; ++e0) {
                const int n = n_of(e0, e1, e2);
                if (n == 13 || (direction_mask && !direction_mask[n]))
                    continue;
                const int e[3] = {e0, e1, e2};
                region r;
                r.count = 1;
                for (int d = 0; d < 3; ++d) {
                    r.lo[d] = lo_outside(desc[d], e[d]);
                    r.len[d] = hi_outside(desc[d], e[d]) - r.lo[d] + 1;
                    r.count *= r.len[d] > 0 ? r.len[d] : 0;
                }
                if (r.count == 0)
                    continue;
                const int sg = a.t.n_seg++;
                a.t.dir[sg] = n;
                a.t.r[sg] = r;
                a.t.buf[sg] = nullptr;
                a.t.flag[sg] = nullptr;
                a.t.chunks_per_field[sg] = (int)((r.count + kChunk - 1) / kChunk);
                a.t.chunk_start[sg + 1] = a.t.chunk_start[sg] + a.t.chunks_per_field[sg] * n_dst;
            }
    if (a.t.n_seg == 0)
        return GTB_OK;
    for (int f = 0; f < n_dst; ++f)
        a.fields[f] = static_cast<char *>(fields[f]);
    a.src = kind == GTB_BC_COPY ? static_cast<const char *>(fields[n_fields - 1]) : nullptr;
    a.s1 = desc[0].total;
    a.s2 = (int64_t)desc[0].total * desc[1].total;
    a.n_fields = n_dst;
    a.fill_bits = value_bits(value, elem_size);
    const int grid = xfer_grid(a.t.chunk_start[a.t.n_seg]);
    cudaStream_t st = as_stream(stream);
    if (kind == GTB_BC_COPY) {
        if (elem_size == 8)
            bc_kernel<uint64_t, true><<<grid, kThreads, 0, st>>>(a);
        else
            bc_kernel<uint32_t, true><<<grid, kThreads, 0, st>>>(a);
    } else {
        if (elem_size == 8)
            bc_kernel<uint64_t, false><<<grid, kThreads, 0, st>>>(a);
        else
            bc_kernel<uint32_t, false><<<grid, kThreads, 0, st>>>(a);
    }
    count_launch();
    return check_launch("boundary apply");
}

GTB_API int gtb_halo_set_boundary(gtb_halo *h, int kind, double value) {
    if (!h)
        return fail(GTB_ERR_ARG, "gtb_halo_set_boundary: null handle");
    if (kind != -1 && kind != GTB_BC_VALUE)
        return fail(GTB_ERR_ARG, "gtb_halo_set_boundary: only GTB_BC_VALUE (or -1 = none) can be fused into the unpack");
    h->bc_kind = kind;
    h->bc_bits = value_bits(value, h->es);
    return GTB_OK;
}

// ------------------------------------------------------------------------------------------- device-side gates
GTB_API void *gtb_halo_unpacked_flag(gtb_halo *h) { return h ? h->d_unpacked : nullptr; }

GTB_API int gtb_halo_gate(gtb_halo *h, const void *counter, uint64_t value) {
    if (!h)
        return fail(GTB_ERR_ARG, "gtb_halo_gate: null handle");
    h->gate_counter = static_cast<const unsigned long long *>(counter);
    h->gate_value = value;
    return GTB_OK;
}

GTB_API uint64_t gtb_halo_epoch(const gtb_halo *h) { return h ? h->epoch : 0; }
