// halo_device.cuh -- device side of the halo exchange: message geometry, arrival flags and the chunk mover of the
// transfer kernels of halo.cu.
#pragma once

#include "common.cuh"

namespace gtb {
    namespace halo_dev {

    constexpr int kMaxFields = 16; // per launch; more fields are handled by looping launches
    constexpr int kThreads = 256;  // block size of the stand-alone transfer kernels
    constexpr int kItems = 8;

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
        for (unsigned spins = 0;; ++spins) {
            uint64_t v; // poll relaxed, acquire once at the end
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= epoch) {
                asm volatile("fence.acq_rel.sys;" ::: "memory");
                return true;
            }
            // the error word lives in HOST memory (a read crosses PCIe): look at it once in a while only
            if (timeout > 0 && (spins & 0xfffu) == 0xfffu &&
                (clock64() - t0 > timeout || *reinterpret_cast<volatile int *>(error))) {
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
        E *__restrict__ buf, const bool fill, const uint64_t fill_bits, const int tid = threadIdx.x) {
        constexpr int kBatch = sizeof(E) >= 16 && ITEMS > 4 ? ITEMS / 2 : ITEMS;
        const uint32_t c = (uint32_t)THREADS / (uint32_t)l0;
        const int d0 = (int)((uint32_t)THREADS - c * (uint32_t)l0), d1 = (int)(c % (uint32_t)l1), d2 = (int)(c / (uint32_t)l1);
        uint32_t e = first + (uint32_t)tid;
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


    } // namespace halo_dev

} // namespace gtb
