/*
 * gtb200/gcl/halo_exchange.hpp -- the gcl halo-exchange interface on top of libgtb200.so (include/gtb200.h).
 *
 *     #include <gtb200/gcl/halo_exchange.hpp>
 *     gtb200::gcl::proc_grid grid({2, 4, 1}, {false, false, false}, rank);
 *     gtb200::gcl::halo_exchange_dynamic_ut<gtb200::gcl::layout_map<2, 1, 0>, gtb200::gcl::layout_map<0, 1, 2>, double>
 *         he({false, false, false}, grid, channel);
 *     he.add_halo<0>(2, 2, 2, ni + 1, ni_total); he.add_halo<1>(...); he.add_halo<2>(0, 0, 0, nk - 1, nk);
 *     he.setup(3);
 *     he.pack(a, b, c); he.exchange(); he.unpack(a, b, c);
 *
 * Mirrors gridtools::gcl::halo_exchange_dynamic_ut<DataLayout, ProcLayout, T, gpu> (gcl/halo_exchange.hpp:163-306):
 * same template parameters (DataLayout is a GridTools layout_map: the dimension with the LARGEST value has stride 1,
 * storage/gpu.hpp uses layout_map<2,1,0>; ProcLayout maps data dimensions to process-grid dimensions), same member
 * functions with the same meaning, same halo_descriptor semantics (common/halo_descriptor.hpp:44-227).  Any type with
 * a static `at(int)` works as a layout, gridtools::layout_map included.
 *
 * What differs, because there is no MPI on the data path:
 *  - `comm` is a gtb200::gcl::proc_grid (the Cartesian grid of gcl/low_level/proc_grids_3D.hpp:34-245: row-major
 *    ranks like MPI_Cart_create, proc(i,j,k) with periodicity :179-211) plus a CHANNEL, a functor
 *        void(const void *mine, void *all, std::size_t bytes)      // all = size * bytes, rank-major
 *    that all-gathers one 512-byte blob per rank.  It is used once, in setup(), to exchange CUDA IPC handles
 *    (MPI_Allgather, torch.distributed, a directory on a shared file system: file_channel below).
 *  - pack() gathers AND pushes: one kernel writes every message straight into the neighbours' receive buffers over
 *    NVLink and raises their arrival flags; exchange()/start_exchange()/wait() have nothing left to do on the host;
 *    unpack() waits for the flags on the device and scatters.  Two launches per exchange, no host synchronisation
 *    (the reference: up to 12 x n_fields launches, cudaDeviceSynchronize, 2 x 26 MPI calls,
 *    gcl/high_level/descriptors_manual_gpu.hpp:342-514, gcl/low_level/Halo_Exchange_3D.hpp:546-931).
 *  - Errors: a non-zero status of the C ABI is thrown as std::runtime_error (the reference asserts or hangs in
 *    MPI_Wait); check() reports a neighbour that never delivered.
 */
#pragma once

#include <array>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../gtb200.h"

namespace gtb200 {
    namespace gcl {

        /// GridTools layout_map semantics: at(d) = rank of dimension d, the largest value is the unit-stride dimension.
        template <int I0, int I1, int I2>
        struct layout_map {
            static constexpr int at(int d) { return d == 0 ? I0 : (d == 1 ? I1 : I2); }
        };

        /// common/halo_descriptor.hpp:44-227
        struct halo_descriptor {
            int minus, plus, begin, end, total_length;
        };

        inline void check(int status, const char *what) {
            if (status != GTB_OK)
                throw std::runtime_error(std::string(what) + ": " + gtb_last_error());
        }

        /// 3-d Cartesian process grid (gcl/low_level/proc_grids_3D.hpp:34-245), row-major ranks like MPI_Cart_create.
        class proc_grid {
            std::array<int, 3> m_dims, m_coords;
            std::array<bool, 3> m_periodic;
            int m_rank;

          public:
            proc_grid(std::array<int, 3> dims, std::array<bool, 3> periodic, int rank)
                : m_dims(dims), m_periodic(periodic), m_rank(rank) {
                if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || rank < 0 || rank >= size())
                    throw std::invalid_argument("gtb200::gcl::proc_grid: rank outside the process grid");
                m_coords = {rank / (dims[1] * dims[2]), (rank / dims[2]) % dims[1], rank % dims[2]};
            }
            int size() const { return m_dims[0] * m_dims[1] * m_dims[2]; }
            int rank() const { return m_rank; }
            void dims(int &i, int &j, int &k) const { i = m_dims[0], j = m_dims[1], k = m_dims[2]; }
            void coords(int &i, int &j, int &k) const { i = m_coords[0], j = m_coords[1], k = m_coords[2]; }
            std::array<int, 3> const &dims() const { return m_dims; }
            std::array<int, 3> const &coords() const { return m_coords; }
            std::array<bool, 3> const &periodic() const { return m_periodic; }
            /// Rank of the process at offset (i, j, k) from this one, -1 outside a non-periodic border (:179-211).
            int proc(int i, int j, int k) const {
                int c[3] = {m_coords[0] + i, m_coords[1] + j, m_coords[2] + k};
                for (int d = 0; d < 3; ++d) {
                    if (m_periodic[d])
                        c[d] = ((c[d] % m_dims[d]) + m_dims[d]) % m_dims[d];
                    else if (c[d] < 0 || c[d] >= m_dims[d])
                        return -1;
                }
                return (c[0] * m_dims[1] + c[1]) * m_dims[2] + c[2];
            }
            /// Balanced factorisation over the first `ndims` dimensions (MPI_Dims_create), larger factor on the later
            /// dimension: with i the unit-stride axis J is split first (J faces are contiguous slabs).
            static std::array<int, 3> dims_create(int nranks, int ndims = 2) {
                std::array<int, 3> d = {1, 1, 1};
                std::vector<int> factors;
                for (int n = nranks, f = 2; n > 1; ++f)
                    while (n % f == 0) {
                        factors.push_back(f);
                        n /= f;
                    }
                for (auto it = factors.rbegin(); it != factors.rend(); ++it) {
                    int best = 0;
                    for (int x = 1; x < ndims; ++x)
                        if (d[x] <= d[best])
                            best = x;
                    d[best] *= *it;
                }
                for (int a = 0; a < ndims; ++a) // sort ascending
                    for (int b = a + 1; b < ndims; ++b)
                        if (d[b] < d[a])
                            std::swap(d[a], d[b]);
                return d;
            }
        };

        /// The out-of-band channel: all-gather `bytes` bytes per rank (rank-major in `all`).
        using channel_t = std::function<void(const void *mine, void *all, std::size_t bytes)>;

        /// A channel over a directory every rank can see (ranks = processes on one box, no MPI): rank r writes
        /// <dir>/<tag>.<r> atomically and polls for the files of the others.
        inline channel_t file_channel(std::string dir, std::string tag, int rank, int size, double timeout_s = 120) {
            return [=](const void *mine, void *all, std::size_t bytes) {
                auto name = [&](int r) { return dir + "/" + tag + "." + std::to_string(r); };
                {
                    std::string tmp = name(rank) + ".tmp";
                    std::ofstream f(tmp, std::ios::binary);
                    f.write(static_cast<const char *>(mine), (std::streamsize)bytes);
                    f.close();
                    if (!f || std::rename(tmp.c_str(), name(rank).c_str()) != 0)
                        throw std::runtime_error("gtb200::gcl::file_channel: cannot write " + name(rank));
                }
                auto t0 = std::chrono::steady_clock::now();
                for (int r = 0; r < size; ++r) {
                    for (;;) {
                        std::ifstream f(name(r), std::ios::binary);
                        if (f && f.read(static_cast<char *>(all) + r * bytes, (std::streamsize)bytes))
                            break;
                        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
                            throw std::runtime_error("gtb200::gcl::file_channel: rank " + std::to_string(r) + " never showed up");
                        std::this_thread::sleep_for(std::chrono::milliseconds(2));
                    }
                }
            };
        }

        template <class DataLayout, class ProcLayout, class T>
        class halo_exchange_dynamic_ut {
            static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- and 8-byte elements");
            // increasing-stride position of user dimension d (the reference's reverse_map, gcl/halo_exchange.hpp:166)
            static constexpr int storage_dim(int d) { return 2 - DataLayout::at(d); }

            proc_grid m_grid;
            channel_t m_channel;
            gtb_halo_desc m_desc[3]; // storage order
            bool m_have[3] = {false, false, false};
            gtb_halo *m_h = nullptr;
            void *m_stream = nullptr;
            int m_max_fields = 0;

            halo_exchange_dynamic_ut(halo_exchange_dynamic_ut const &) = delete;
            halo_exchange_dynamic_ut &operator=(halo_exchange_dynamic_ut const &) = delete;

            void **ptrs(std::vector<T *> const &f) { return reinterpret_cast<void **>(const_cast<T **>(f.data())); }

          public:
            using grid_type = proc_grid;
            static constexpr int DIMS = 3;

            /// periodicity in the order of the DATA dimensions (gcl/halo_exchange.hpp:202 permutes it the same way)
            halo_exchange_dynamic_ut(std::array<bool, 3> periodicity, proc_grid const &grid, channel_t channel)
                : m_grid(grid), m_channel(std::move(channel)) {
                std::array<bool, 3> per = {false, false, false};
                for (int d = 0; d < 3; ++d)
                    per[ProcLayout::at(d)] = periodicity[d];
                if (per != grid.periodic())
                    throw std::invalid_argument("gtb200::gcl::halo_exchange_dynamic_ut: periodicity does not match the process grid");
            }
            ~halo_exchange_dynamic_ut() {
                if (m_h)
                    gtb_halo_destroy(m_h);
            }

            /// add_halo<DI>(minus, plus, begin, end, total) -- DI in the logical order of the application (:235)
            template <int DI>
            void add_halo(int minus, int plus, int begin, int end, int t_len) {
                static_assert(DI >= 0 && DI < 3, "dimension index");
                m_desc[storage_dim(DI)] = gtb_halo_desc{minus, plus, begin, end, t_len};
                m_have[storage_dim(DI)] = true;
            }
            template <int DI>
            void add_halo(halo_descriptor const &h) {
                add_halo<DI>(h.minus, h.plus, h.begin, h.end, h.total_length);
            }

            /// Kernels are enqueued on this cudaStream_t (default: the legacy default stream, like the reference).
            void set_stream(void *cuda_stream) { m_stream = cuda_stream; }

            /// setup(max_fields) (:216): buffers, neighbour table, exchange of the IPC handles through the channel.
            void setup(int max_fields_n) {
                if (!(m_have[0] && m_have[1] && m_have[2]))
                    throw std::logic_error("gtb200::gcl: setup() before add_halo() for all three dimensions");
                if (m_h)
                    throw std::logic_error("gtb200::gcl: setup() called twice");
                int nbr[27];
                for (int n = 0; n < 27; ++n) {
                    const int es[3] = {n % 3 - 1, (n / 3) % 3 - 1, n / 9 - 1}; // offsets in storage order
                    int off[3]; // process dimension i moves with user dimension ProcLayout::at(i) (descriptors.hpp:497-499)
                    for (int i = 0; i < 3; ++i)
                        off[i] = es[storage_dim(ProcLayout::at(i))];
                    nbr[n] = n == 13 ? -1 : m_grid.proc(off[0], off[1], off[2]);
                }
                check(gtb_halo_create(m_desc, nbr, m_grid.rank(), max_fields_n, (int)sizeof(T), &m_h), "gtb_halo_create");
                m_max_fields = max_fields_n;
                std::vector<char> mine(GTB_HALO_BLOB_BYTES), all((std::size_t)m_grid.size() * GTB_HALO_BLOB_BYTES);
                check(gtb_halo_export(m_h, mine.data()), "gtb_halo_export");
                m_channel(mine.data(), all.data(), GTB_HALO_BLOB_BYTES);
                const void *blobs[27];
                for (int n = 0; n < 27; ++n)
                    blobs[n] = nbr[n] >= 0 ? all.data() + (std::size_t)nbr[n] * GTB_HALO_BLOB_BYTES : nullptr;
                check(gtb_halo_connect(m_h, blobs), "gtb_halo_connect");
            }

            /// pack(fields) (:250,:269): gather + NVLink push + signal, one launch.
            void pack(std::vector<T *> const &fields) {
                int lost = 0; // a wait of an earlier exchange that gave up (halo.timeout_ms): a host memory read
                check(gtb_halo_poll_error(m_h, &lost), "gtb_halo_poll_error");
                if (lost)
                    throw std::runtime_error("gtb200::gcl: the message from direction " + std::to_string(lost - 1) +
                                             " never arrived; the halos of that exchange are stale");
                check(gtb_halo_pack_send(m_h, ptrs(fields), (int)fields.size(), m_stream), "gtb_halo_pack_send");
            }
            template <class... Fields>
            void pack(const Fields *...fields) {
                pack(std::vector<T *>{const_cast<T *>(static_cast<const T *>(fields))...});
            }

            /// exchange() = start_exchange() + wait() (:284-304): the messages are already on their way.
            void exchange() {}
            void post_receives() {}
            void do_sends() {}
            void start_exchange() {}
            void wait() {}

            /// unpack(fields) (:260,:276): device-side wait for the neighbours' flags + scatter, one launch.
            void unpack(std::vector<T *> const &fields) {
                check(gtb_halo_wait_unpack(m_h, ptrs(fields), (int)fields.size(), m_stream), "gtb_halo_wait_unpack");
                check(gtb_halo_next_epoch(m_h), "gtb_halo_next_epoch");
            }
            template <class... Fields>
            void unpack(Fields *...fields) {
                unpack(std::vector<T *>{static_cast<T *>(fields)...});
            }

            /// Synchronises the device; 0 if every message arrived, 1 + direction of one that did not.
            int check_arrivals() {
                int code = 0;
                check(gtb_halo_error(m_h, &code), "gtb_halo_error");
                return code;
            }

            grid_type const &comm() const { return m_grid; }
        };

    } // namespace gcl
} // namespace gtb200
