/*
 * gtb200/gcl/b200.hpp -- the B200 halo exchange behind gridtools::gcl's OWN class templates.
 *
 *     #include <gridtools/gcl/halo_exchange.hpp>
 *     #include <gtb200/gcl/b200.hpp>                       // adds the arch tag gridtools::gcl::b200
 *     using pattern_t = gridtools::gcl::halo_exchange_dynamic_ut<layout_map<0, 1, 2>, layout_map<0, 1, 2>, double,
 *                                                               gridtools::gcl::b200>;   // was: gcl::gpu
 *     pattern_t he(gridtools::gcl::boollist<3>(false, false, false), CartComm);          // unchanged ctor (:202)
 *     he.add_halo<0>(...); he.setup(3); he.pack(a, b, c); he.exchange(); he.unpack(a, b, c);
 *
 * The reference's halo_exchange_dynamic_ut<DataLayout, ProcLayout, T, Arch> (gcl/halo_exchange.hpp:163-306) is a thin
 * shell around hndlr_dynamic_ut<T, grid, pattern, proc_layout, Arch> (:185) -- the reference's GPU path is exactly
 * such a handler specialisation for gcl::gpu (gcl/high_level/descriptors_manual_gpu.hpp:83-515).  This header adds the
 * arch tag `gridtools::gcl::b200` next to gcl::cpu / gcl::gpu (gcl/low_level/arch.hpp:28,32) and the handler
 * specialisation for it, so the reference's class -- constructor (periodicity, MPI_Comm), add_halo<D>, setup, pack /
 * unpack (variadic and vector), exchange / post_receives / do_sends / start_exchange / wait, comm() -- works
 * unchanged with only the arch tag replaced.  Likewise halo_exchange_generic<ProcLayout, gcl::b200> +
 * field_on_the_fly (gcl/halo_exchange.hpp:335-513, gcl/high_level/descriptor_generic_manual.hpp:370-796).
 *
 * What the handler does differently from the gpu one: MPI is used ONCE, in setup(), to all-gather one 512-byte blob
 * per rank (CUDA IPC handle of the receive arena) over the Cartesian communicator.  After that pack() is one launch
 * that gathers every field for every neighbour and stores the messages straight into the neighbours' receive
 * buffers over NVLink, unpack() one launch that acquires the arrival flags on the device and scatters; exchange()
 * and its split-phase parts have nothing left to do (no MPI_Isend / MPI_Irecv / MPI_Wait, no cudaDeviceSynchronize).
 * Ranks must share one NVLink/NVSwitch box (one process per GPU, or threads of one process).
 *
 * Synchronisation: by default pack() and unpack() behave like the reference's GPU handler, which ends both with
 * cudaDeviceSynchronize (descriptors_manual_gpu.hpp:400,431,486,514): the kernels run on a private non-blocking stream
 * ordered after the legacy default stream, and unpack() returns when the halos are in place.  set_stream(s) switches
 * to fully asynchronous operation on the caller's stream (no host synchronisation at all; the caller orders the
 * stream against its stencils), which is what overlaps an exchange with computation.
 *
 * Element types of any size that is a multiple of 4 bytes travel as 4- or 8-byte words (the reference's own test
 * exchanges array<int, 4>).  Failures of the C ABI are thrown as std::runtime_error.
 */
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include <mpi.h>

#include <gridtools/common/array.hpp>
#include <gridtools/common/halo_descriptor.hpp>
#include <gridtools/gcl/halo_exchange.hpp>

#include "../../gtb200.h"
#include "arch.hpp"

namespace gridtools {
    namespace gcl {
        namespace b200_impl_ {
            inline void check(int status, const char *what) {
                if (status != GTB_OK)
                    throw std::runtime_error(std::string(what) + ": " + gtb_last_error());
            }

            /// A device-side wait that gave up (option halo.timeout_ms) left its halo untouched and set the object's
            /// error word; the reference would hang in MPI_Wait.  Cheap (a host memory read): called in every pack / unpack.
            inline void throw_if_lost(gtb_halo *h) {
                int code = 0;
                check(gtb_halo_poll_error(h, &code), "gtb_halo_poll_error");
                if (code)
                    throw std::runtime_error("gcl::b200: the message from direction " + std::to_string(code - 1) +
                                             " never arrived (halo.timeout_ms); the halos of that exchange are stale");
            }

            /// default: a private non-blocking stream, ordered after the legacy default stream at pack(), host-synchronised
            /// at the end of unpack() (the reference's blocking semantics); a user stream: nothing of that
            class stream_policy {
                void *m_own = nullptr, *m_user = nullptr;

              public:
                stream_policy() = default;
                stream_policy(stream_policy const &) = delete;
                ~stream_policy() {
                    if (m_own)
                        gtb_stream_destroy(m_own);
                }
                void set(void *user) { m_user = user; }
                void *get() {
                    if (m_user)
                        return m_user;
                    if (!m_own)
                        check(gtb_stream_create(&m_own, 1), "gtb_stream_create");
                    return m_own;
                }
                void *before_pack() {
                    void *s = get();
                    if (!m_user)
                        check(gtb_stream_after_default(s), "gtb_stream_after_default");
                    return s;
                }
                /// true if the host waited for the unpack (default mode): its outcome is known now
                bool after_unpack() {
                    if (!m_user)
                        check(gtb_stream_synchronize(m_own), "gtb_stream_synchronize");
                    return !m_user;
                }
            };

            /// words a T is moved as, and how many of them
            template <class T>
            struct words {
                static_assert(sizeof(T) % 4 == 0, "gcl::b200 moves elements as 4- or 8-byte words");
                static constexpr int size = sizeof(T) % 8 == 0 ? 8 : 4;
                static constexpr int count = sizeof(T) / size;
            };

            inline gtb_halo_desc scaled(halo_descriptor const &h, int f) {
                return gtb_halo_desc{(int)h.minus() * f, (int)h.plus() * f, (int)h.begin() * f, (int)h.end() * f + f - 1,
                    (int)h.total_length() * f};
            }

            /// neighbour table in storage-direction numbering n = (e0+1) + 3 (e1+1) + 9 (e2+1); the process-grid offsets
            /// of a storage direction are taken exactly like the reference's handlers do (descriptors.hpp:497-499)
            template <class ProcLayout, class Grid>
            void neighbours(Grid const &grid, int nbr[27]) {
                for (int kk = -1; kk <= 1; ++kk)
                    for (int jj = -1; jj <= 1; ++jj)
                        for (int ii = -1; ii <= 1; ++ii) {
                            const int n = (ii + 1) + 3 * (jj + 1) + 9 * (kk + 1);
                            nbr[n] = n == 13 ? -1
                                             : grid.proc(nth<ProcLayout, 0>(ii, jj, kk), nth<ProcLayout, 1>(ii, jj, kk),
                                                   nth<ProcLayout, 2>(ii, jj, kk));
                        }
            }

            /// setup(): create the device object, all-gather the connection blobs over the Cartesian communicator
            template <class Grid>
            gtb_halo *create_and_connect(
                Grid const &grid, const gtb_halo_desc desc[3], const int nbr[27], int max_fields, int word_size) {
                gtb_halo *h = nullptr;
                check(gtb_halo_create(desc, nbr, grid.pid(), max_fields, word_size, &h), "gtb_halo_create");
                int size = 0;
                MPI_Comm_size(grid.communicator(), &size);
                std::vector<char> mine(GTB_HALO_BLOB_BYTES), all((std::size_t)size * GTB_HALO_BLOB_BYTES);
                check(gtb_halo_export(h, mine.data()), "gtb_halo_export");
                MPI_Allgather(mine.data(), GTB_HALO_BLOB_BYTES, MPI_CHAR, all.data(), GTB_HALO_BLOB_BYTES, MPI_CHAR,
                    grid.communicator());
                const void *blobs[27];
                for (int n = 0; n < 27; ++n)
                    blobs[n] = nbr[n] >= 0 ? all.data() + (std::size_t)nbr[n] * GTB_HALO_BLOB_BYTES : nullptr;
                check(gtb_halo_connect(h, blobs), "gtb_halo_connect");
                return h;
            }
        } // namespace b200_impl_

        /** hndlr_dynamic_ut for gcl::b200: what halo_exchange_dynamic_ut<..., gcl::b200> delegates to
            (the counterpart of descriptors_manual_gpu.hpp:83-515). */
        template <typename DataType, typename HaloExch, typename proc_layout, template <int Ndim> class GridType>
        class hndlr_dynamic_ut<DataType, GridType<3>, HaloExch, proc_layout, b200> : public descriptor_base<HaloExch> {
            using words = b200_impl_::words<DataType>;

            gtb_halo *m_h = nullptr;
            b200_impl_::stream_policy m_stream;

            hndlr_dynamic_ut(hndlr_dynamic_ut const &) = delete;
            hndlr_dynamic_ut(hndlr_dynamic_ut &&) = delete;

            void do_pack(void *const *fields, int n) {
                if (!m_h)
                    throw std::logic_error("gcl::b200: pack() before setup()");
                b200_impl_::throw_if_lost(m_h);
                b200_impl_::check(gtb_halo_pack_send(m_h, fields, n, m_stream.before_pack()), "gtb_halo_pack_send");
            }
            void do_unpack(void *const *fields, int n) {
                if (!m_h)
                    throw std::logic_error("gcl::b200: unpack() before setup()");
                b200_impl_::check(gtb_halo_wait_unpack(m_h, fields, n, m_stream.get()), "gtb_halo_wait_unpack");
                b200_impl_::check(gtb_halo_next_epoch(m_h), "gtb_halo_next_epoch");
                if (m_stream.after_unpack())
                    b200_impl_::throw_if_lost(m_h);
            }

          public:
            typedef b200 arch_type;
            typedef descriptor_base<HaloExch> base_type;
            typedef typename base_type::pattern_type pattern_type;
            typedef typename pattern_type::grid_type grid_type;

            /// halo descriptors in increasing-stride order, filled by halo_exchange_dynamic_ut::add_halo (:235-242)
            empty_field_no_dt halo;

            explicit hndlr_dynamic_ut(typename grid_type::period_type const &c, MPI_Comm const &comm)
                : base_type(c, comm), halo() {}
            explicit hndlr_dynamic_ut(grid_type const &g) : base_type(g), halo() {}
            ~hndlr_dynamic_ut() {
                if (m_h)
                    gtb_halo_destroy(m_h);
            }

            /// setup(max_fields_n) (gcl/halo_exchange.hpp:216): buffers, neighbour table, connection
            void setup(int max_fields_n) {
                if (m_h)
                    throw std::logic_error("gcl::b200: setup() called twice");
                gtb_halo_desc desc[3];
                for (int d = 0; d < 3; ++d)
                    desc[d] = b200_impl_::scaled(halo.halos[d], d == 0 ? words::count : 1);
                int nbr[27];
                b200_impl_::neighbours<proc_layout>(this->comm(), nbr);
                m_h = b200_impl_::create_and_connect(this->comm(), desc, nbr, max_fields_n, words::size);
            }

            /// asynchronous operation on the caller's cudaStream_t (see the header comment); nullptr = the default again
            void set_stream(void *cuda_stream) { m_stream.set(cuda_stream); }

            template <typename... FIELDS>
            void pack(const FIELDS *..._fields) {
                void *f[] = {const_cast<void *>(static_cast<const void *>(_fields))...};
                do_pack(f, (int)sizeof...(FIELDS));
            }
            template <typename... FIELDS>
            void unpack(FIELDS *..._fields) {
                void *f[] = {static_cast<void *>(_fields)...};
                do_unpack(f, (int)sizeof...(FIELDS));
            }
            void pack(std::vector<DataType *> const &fields) {
                do_pack(reinterpret_cast<void *const *>(fields.data()), (int)fields.size());
            }
            void unpack(std::vector<DataType *> const &fields) {
                do_unpack(reinterpret_cast<void *const *>(fields.data()), (int)fields.size());
            }

            /// the messages left with pack(); arrival is awaited on the device inside unpack()
            void exchange() {}
            void post_receives() {}
            void do_sends() {}
            void start_exchange() {}
            void wait() {}

            /// Synchronises the device; 0 if every message arrived, 1 + direction of one that did not.
            int check_arrivals() {
                int code = 0;
                b200_impl_::check(gtb_halo_error(m_h, &code), "gtb_halo_error");
                return code;
            }

            pattern_type const &pattern() const { return base_type::pattern(); }
        };

        /** hndlr_generic for gcl::b200 (counterpart of descriptor_generic_manual.hpp:370-796): every field_on_the_fly
            brings its own halo descriptors and element type; all fields of a pack() travel in ONE message per
            neighbour, packed by one launch and unpacked by one launch. */
        template <typename HaloExch, typename proc_layout_abs>
        class hndlr_generic<HaloExch, proc_layout_abs, b200> : public descriptor_base<HaloExch> {
            gtb_halo *m_h = nullptr;
            mutable b200_impl_::stream_policy m_stream;
            int m_max_fields = 0;

            hndlr_generic(hndlr_generic const &) = delete;

            template <class Fotf>
            static gtb_halo_field make(Fotf const &f) {
                using T = typename Fotf::value_type;
                using words = b200_impl_::words<T>;
                gtb_halo_field g;
                g.ptr = const_cast<void *>(static_cast<const void *>(f.ptr));
                for (int d = 0; d < 3; ++d)
                    g.desc[d] = b200_impl_::scaled(f.halos[d], d == 0 ? (int)(sizeof(T) / 4) : 1);
                (void)sizeof(words);
                return g;
            }

            void do_pack(std::vector<gtb_halo_field> const &fs) const {
                if (!m_h)
                    throw std::logic_error("gcl::b200: pack() before setup()");
                b200_impl_::throw_if_lost(m_h);
                b200_impl_::check(gtb_halo_generic_pack_send(m_h, fs.data(), (int)fs.size(), m_stream.before_pack()),
                    "gtb_halo_generic_pack_send");
            }
            void do_unpack(std::vector<gtb_halo_field> const &fs) const {
                if (!m_h)
                    throw std::logic_error("gcl::b200: unpack() before setup()");
                b200_impl_::check(gtb_halo_generic_wait_unpack(m_h, fs.data(), (int)fs.size(), m_stream.get()),
                    "gtb_halo_generic_wait_unpack");
                b200_impl_::check(gtb_halo_next_epoch(m_h), "gtb_halo_next_epoch");
                if (m_stream.after_unpack())
                    b200_impl_::throw_if_lost(m_h);
            }

          public:
            typedef descriptor_base<HaloExch> base_type;
            typedef typename base_type::pattern_type pattern_type;
            typedef typename pattern_type::grid_type grid_type;

            explicit hndlr_generic(grid_type const &g) : base_type(g) {}
            ~hndlr_generic() {
                if (m_h)
                    gtb_halo_destroy(m_h);
            }

            /** setup(max_fields_n, halo_example, typesize) (gcl/halo_exchange.hpp:389-393): buffers sized for
                max_fields_n fields of the example's halo regions with typesize-byte elements */
            template <typename DataType, typename f_data_layout, template <typename> class traits>
            void setup(int max_fields_n, field_on_the_fly<DataType, f_data_layout, traits> const &halo_example,
                int typesize) {
                if (m_h)
                    throw std::logic_error("gcl::b200: setup() called twice");
                if (typesize % 4)
                    throw std::invalid_argument("gcl::b200: element sizes must be multiples of 4 bytes");
                using proc_layout = layout_transform<typename field_on_the_fly<DataType, f_data_layout, traits>::inner_layoutmap,
                    proc_layout_abs>;
                gtb_halo_desc desc[3];
                for (int d = 0; d < 3; ++d)
                    desc[d] = b200_impl_::scaled(halo_example.halos[d], d == 0 ? typesize / 4 : 1);
                int nbr[27];
                b200_impl_::neighbours<proc_layout>(this->comm(), nbr);
                m_h = b200_impl_::create_and_connect(this->comm(), desc, nbr, max_fields_n, 4);
                m_max_fields = max_fields_n;
            }

            void set_stream(void *cuda_stream) { m_stream.set(cuda_stream); }

            template <typename... FIELDS>
            void pack(const FIELDS &..._fields) const {
                do_pack(std::vector<gtb_halo_field>{make(_fields)...});
            }
            template <typename... FIELDS>
            void unpack(const FIELDS &..._fields) const {
                do_unpack(std::vector<gtb_halo_field>{make(_fields)...});
            }
            template <typename T1, typename T2, template <typename> class T3>
            void pack(std::vector<field_on_the_fly<T1, T2, T3>> const &fields) {
                std::vector<gtb_halo_field> fs;
                for (auto const &f : fields)
                    fs.push_back(make(f));
                do_pack(fs);
            }
            template <typename T1, typename T2, template <typename> class T3>
            void unpack(std::vector<field_on_the_fly<T1, T2, T3>> const &fields) {
                std::vector<gtb_halo_field> fs;
                for (auto const &f : fields)
                    fs.push_back(make(f));
                do_unpack(fs);
            }

            void exchange() {}
            void post_receives() {}
            void do_sends() {}
            void start_exchange() {}
            void wait() {}

            int check_arrivals() {
                int code = 0;
                b200_impl_::check(gtb_halo_error(m_h, &code), "gtb_halo_error");
                return code;
            }

            pattern_type const &pattern() const { return base_type::pattern(); }
        };

        /// halo_exchange_generic<ProcLayout, gcl::b200> (next to the cpu / gpu specialisations, halo_exchange.hpp:470-513)
        template <typename layout2proc_map>
        class halo_exchange_generic<layout2proc_map, b200> : public halo_exchange_generic_base<layout2proc_map, b200> {
            typedef halo_exchange_generic_base<layout2proc_map, b200> base_type;

          public:
            typedef typename base_type::grid_type grid_type;
            typedef typename base_type::pattern_type pattern_type;

            template <typename DT>
            struct traits {
                static const int I = 3;
                typedef empty_field_no_dt base_field; // host-side descriptors only
            };

            explicit halo_exchange_generic(typename grid_type::period_type const &c, MPI_Comm comm) : base_type(c, comm) {}
            explicit halo_exchange_generic(grid_type const &g) : base_type(g) {}
        };
    } // namespace gcl
} // namespace gridtools
