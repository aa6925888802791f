// halo_device.cuh -- device side of the halo exchange that more than one translation unit needs: the transfer kernels
// of halo.cu and the COMMUNICATION CTAs that a stencil launch can carry (gtb_halo_attach, include/gtb200.h): a few
// extra CTAs of the stencil kernel's own grid run a complete exchange (pack -> NVLink stores -> flags -> wait ->
// unpack) beside the CTAs that compute, so that a time step is ONE launch with no stream events around it.
#pragma once

#include "common.cuh"

namespace gtb {
    namespace halo_dev {

    constexpr int kMaxFields = 16; // per launch; more fields are handled by looping launches
    constexpr int kThreads = 256;  // block size of the stand-alone transfer kernels
    constexpr int kItems = 8;
    constexpr int kAttachedItems = 16; // elements per thread and chunk of an attached exchange: chunk = 16 x CTA size

    struct region {
        int lo[3];
        int len[3];
        int64_t count; // 0: no such neighbour
    };


    // Fused synchronisation of a transfer launch.  mode 1 (pack towards peers): the last block of the launch raises
    // the neighbours' flags once every block has made its stores visible system-wide.  mode 2 (unpack): a block
    // acquires the flag of a direction before it touches that direction's message.
    struct sync_args {
        uint64_t epoch;
        unsigned *counter; // zero between launches
        int *error;
        long long timeout_cycles;
        int mode; // 0 none, 1 signal after pack, 2 wait before unpack
        const unsigned long long *gate; // unpack: wait for *gate >= gate_value before scattering (stencil still reads the halos)
        unsigned long long gate_value;
        unsigned long long *gate_timeouts;
        unsigned long long *unpacked;   // unpack: the last block stores the epoch here when every block is done
        unsigned *counter2;             // block counter of that, zero between launches
        unsigned long long *trace;      // diagnosis: row of 8 globaltimer stamps of this epoch, or nullptr
    };

    constexpr int kMaxSeg = 26;
    constexpr int kChunk = kThreads * kItems; // elements a block of a stand-alone transfer kernel moves per step

    // One launch moves every field of every active direction ("segment").  The work is a flat list of chunks of
    // kChunk elements, segment-major then field-major, that a SMALL grid walks with a grid stride: the exchange has
    // to run beside a persistent stencil kernel that owns the SMs, so it must not flood the CTA scheduler -- the first
    // version launched 27 x n_fields x ceil(count / 1024) blocks, most of them empty, which filled every thread slot
    // of the chip ahead of the lower-priority stencil launch and serialised the two (profiles/README.md).
    struct seg_table {
        int n_seg;
        int dir[kMaxSeg];             // direction number (0..26) of segment s: reported by a wait that times out
        region r[kMaxSeg];
        char *buf[kMaxSeg];           // message buffer per segment (local or NVLink-mapped)
        uint64_t *flag[kMaxSeg];      // flag to raise (pack) or to wait for (unpack); nullptr: none
        int chunks_per_field[kMaxSeg];
        int chunk_start[kMaxSeg + 1]; // prefix sum of chunks_per_field * n_fields
    };

    struct exchange_args { // pack + signal + wait + unpack in one launch
        seg_table snd, rcv;
        char *fields[kMaxFields];
        int64_t s1, s2;
        int n_fields;
        uint64_t fill_bits;
        sync_args sync;
    };


    // Waits until *flag >= epoch.  Returns false if it gave up: timeout (clock cycles) <= 0 waits for ever, like the
    // MPI_Wait of the reference; with a timeout the caller must NOT unpack the message and the error word (mapped
    // host memory: the host sees it without synchronising) holds 1 + direction.  An error already set makes later
    // waits fail at once instead of stalling every following exchange for the whole timeout as well.
    __device__ __forceinline__ bool wait_flag(const uint64_t *flag, uint64_t epoch, int *error, long long timeout, int n) {
        const long long t0 = clock64();
        for (;;) {
            uint64_t v; // poll relaxed, acquire once at the end
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= epoch) {
                asm volatile("fence.acq_rel.sys;" ::: "memory");
                return true;
            }
            if (timeout > 0 && (clock64() - t0 > timeout || *reinterpret_cast<volatile int *>(error))) {
                if (!*reinterpret_cast<volatile int *>(error)) {
                    *reinterpret_cast<volatile int *>(error) = 1 + n;
                    __threadfence_system();
                }
                return false;
            }
            __nanosleep(100);
        }
    }

    template <class E>
    __device__ __forceinline__ E make_fill(uint64_t bits) {
        return (E)bits;
    }
    template <>
    __device__ __forceinline__ uint4 make_fill<uint4>(uint64_t bits) { // bits = the element pattern repeated to 64 bits
        return make_uint4((unsigned)bits, (unsigned)(bits >> 32), (unsigned)bits, (unsigned)(bits >> 32));
    }

    // One chunk (kChunk elements starting at element `first` of a box region) between a field and a message buffer.
    // The exchange runs on the few SMs a persistent stencil kernel leaves free, so instructions count: positions are
    // 32-bit, a thread locates its first element with two divisions per chunk and then steps kThreads elements at a
    // time with carries (the first version divided 64-bit numbers twice per element: 40 M instructions per exchange,
    // issue-bound on four SMs); the scattered side needs no index array, which keeps the kernel at a register count
    // that lets six blocks share an SM.  Elements move in batches of 64 bytes per thread.
    template <class E, bool PACK, int THREADS = kThreads, int ITEMS = kItems>
    __device__ __forceinline__ void move_chunk(const int lo0, const int lo1, const int lo2, const int l0, const int l1,
        const int64_t s1, const int64_t s2, const uint32_t count, const uint32_t first, E *__restrict__ fld,
        E *__restrict__ buf, const bool fill, const uint64_t fill_bits) {
        constexpr int kBatch = sizeof(E) >= 16 && ITEMS > 4 ? ITEMS / 2 : ITEMS;
        const uint32_t c = (uint32_t)THREADS / (uint32_t)l0;
        const int d0 = (int)((uint32_t)THREADS - c * (uint32_t)l0), d1 = (int)(c % (uint32_t)l1), d2 = (int)(c / (uint32_t)l1);
        uint32_t e = first + threadIdx.x;
        const uint32_t q = e / (uint32_t)l0, q2 = q / (uint32_t)l1;
        int i0 = (int)(e - q * (uint32_t)l0), i1 = (int)(q - q2 * (uint32_t)l1), i2 = (int)q2;
        auto index = [&]() { return (int64_t)(lo0 + i0) + (int64_t)(lo1 + i1) * s1 + (int64_t)(lo2 + i2) * s2; };
        auto advance = [&]() {
            i0 += d0;
            if (i0 >= l0) {
                i0 -= l0;
                ++i1;
            }
            i1 += d1;
            if (i1 >= l1) {
                i1 -= l1;
                ++i2;
            }
            i2 += d2;
        };
#pragma unroll 1
        for (int b = 0; b < ITEMS / kBatch; ++b) {
            E v[kBatch];
            if (PACK) {
#pragma unroll
                for (int it = 0; it < kBatch; ++it) {
                    if (e + (uint32_t)(it * THREADS) < count)
                        v[it] = fld[index()];
                    advance();
                }
#pragma unroll
                for (int it = 0; it < kBatch; ++it)
                    if (e + (uint32_t)(it * THREADS) < count)
                        buf[e + (uint32_t)(it * THREADS)] = v[it];
            } else {
#pragma unroll
                for (int it = 0; it < kBatch; ++it)
                    if (e + (uint32_t)(it * THREADS) < count)
                        v[it] = fill ? make_fill<E>(fill_bits) : buf[e + (uint32_t)(it * THREADS)];
#pragma unroll
                for (int it = 0; it < kBatch; ++it) {
                    if (e + (uint32_t)(it * THREADS) < count)
                        fld[index()] = v[it];
                    advance();
                }
            }
            e += (uint32_t)(kBatch * THREADS);
        }
    }


    // ------------------------------------------------------------------------------------------ communication CTAs
    // What CTA `cta` of `n_cta` communication CTAs does inside a stencil launch: its share of the pack chunks, the
    // hand-shake (the last CTA to finish raises the neighbours' flags), its share of the unpack chunks behind the
    // arrival flags.  A communication CTA only ever waits for communication CTAs of the NEIGHBOURS' launches, never
    // for a CTA of its own launch, and the launchers put these CTAs first in the grid so that they are resident
    // before any stencil CTA.  Chunks are 16 elements per thread: 16 loads in flight per thread (the exchange is
    // latency-bound on the few SMs it gets).
    struct attached_args {
        exchange_args x;
        int n_cta;  // 0: no exchange attached to this launch
        int es;     // bytes per element the tables are expressed in: 4, 8 or 16
        int chunk;  // elements per chunk the tables were built with (16 x block size of the carrying kernel)
    };

    template <class E, bool PACK, int THREADS>
    __device__ __forceinline__ void comm_move(const seg_table &t, char *const *fields, int64_t s1, int64_t s2,
        const sync_args &sy, uint64_t fill_bits, int cta, int n_cta, int *s_ok) {
        constexpr int ITEMS = kAttachedItems;
        constexpr int kAttachedChunk = THREADS * ITEMS;
        const int total = t.chunk_start[t.n_seg];
        unsigned waited = 0, failed = 0;
        int s = 0;
        for (int ch = cta; ch < total; ch += n_cta) {
            while (ch >= t.chunk_start[s + 1])
                ++s;
            if (!PACK && t.flag[s] && !((waited >> s) & 1u)) {
                if (threadIdx.x == 0)
                    *s_ok = wait_flag(t.flag[s], sy.epoch, sy.error, sy.timeout_cycles, t.dir[s]);
                __syncthreads();
                waited |= 1u << s;
                if (!*s_ok)
                    failed |= 1u << s;
                __syncthreads();
            }
            if ((failed >> s) & 1u)
                continue; // the message never arrived: leave the halo alone, the error flag is set
            const region &r = t.r[s];
            const int local = ch - t.chunk_start[s];
            const int f = local / t.chunks_per_field[s];
            const uint32_t first = (uint32_t)(local - f * t.chunks_per_field[s]) * (uint32_t)kAttachedChunk;
            const bool fill = !PACK && t.buf[s] == nullptr;
            E *buf = reinterpret_cast<E *>(t.buf[s]) + (int64_t)f * r.count;
            move_chunk<E, PACK, THREADS, ITEMS>(r.lo[0], r.lo[1], r.lo[2], r.len[0], r.len[1], s1, s2, (uint32_t)r.count,
                first, reinterpret_cast<E *>(fields[f]), buf, fill, fill_bits);
        }
    }

    template <class E, int THREADS>
    __device__ __forceinline__ void comm_cta_typed(const exchange_args &a, int cta, int n_cta, int *s_int) {
        if (a.sync.trace && threadIdx.x == 0 && cta == 0)
            a.sync.trace[0] = ptx::globaltimer();
        comm_move<E, true, THREADS>(a.snd, a.fields, a.s1, a.s2, a.sync, 0, cta, n_cta, s_int);
        __syncthreads(); // this CTA's payload stores, then ONE cumulative system-scope fence
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned done = atomicAdd(a.sync.counter, 1u) + 1u;
            *s_int = done == (unsigned)n_cta;
            if (*s_int)
                *a.sync.counter = 0;
        }
        __syncthreads();
        if (*s_int && threadIdx.x < a.snd.n_seg && a.snd.flag[threadIdx.x]) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.snd.flag[threadIdx.x]), "l"(a.sync.epoch) : "memory");
        }
        __syncthreads();
        if (a.sync.trace && threadIdx.x == 0) {
            atomicMax(a.sync.trace + 1, ptx::globaltimer());
            if (cta == 0)
                a.sync.trace[2] = ptx::globaltimer();
        }
        comm_move<E, false, THREADS>(a.rcv, a.fields, a.s1, a.s2, a.sync, a.fill_bits, cta, n_cta, s_int);
        if (a.sync.trace && threadIdx.x == 0)
            atomicMax(a.sync.trace + 4, ptx::globaltimer());
    }

    // s_int: one int of shared memory
    template <int THREADS>
    __device__ __forceinline__ void comm_cta(const attached_args &a, int cta, int *s_int) {
        if (a.es == 8)
            comm_cta_typed<uint64_t, THREADS>(a.x, cta, a.n_cta, s_int);
        else if (a.es == 4)
            comm_cta_typed<uint32_t, THREADS>(a.x, cta, a.n_cta, s_int);
        else
            comm_cta_typed<uint4, THREADS>(a.x, cta, a.n_cta, s_int);
    }

    } // namespace halo_dev

    // host side (halo.cu): the exchange armed by gtb_halo_attach for the next stencil launch of this thread, if any.
    // Fills `out` (n_cta = 0 when nothing is attached) and advances the epoch of the halo object.
    // cta_threads: block size of the kernel that will carry it (the chunk tables depend on it).
    int take_attached(halo_dev::attached_args &out, int cta_threads);
    // called after every stencil launch: an attached exchange that was not taken runs as launches of its own
    int flush_attached(void *stream);
} // namespace gtb
